"""Stand-in for `joblib` used ONLY by tools/mlprobs_overlay (it shadows the real package because MLProbs' scripts are run
with this directory first on sys.path).  MLProbs' three random-forest classifiers were pickled with scikit-learn 0.21.3
(requirements.txt:142); a current scikit-learn cannot unpickle them (module paths and the tree node dtype changed).  `load`
reads the joblib file with the real joblib's unpickler, maps every sklearn class to a plain placeholder, and returns an
object with the one method the driver calls, `predict` (utils/classifier_c_p_np_aln.py:22-23,
classifier_realign_strategy.py:26-27, classifier_region_min_length.py:26-27), evaluated the way
RandomForestClassifier.predict does: features cast to float32, `x <= threshold` goes left, per-tree class probabilities =
leaf value / its sum, averaged over the trees, arg-max (first maximum) mapped through classes_."""
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


_REAL = None


def _real_numpy_pickle():
    """(joblib.numpy_pickle of the real package, prefix under which the real package's modules sit in sys.modules)."""
    global _REAL
    if _REAL is not None:
        return _REAL
    me = sys.modules.get("joblib")
    if me is not None and os.path.abspath(getattr(me, "__file__", "") or "") != os.path.abspath(__file__):
        _REAL = (importlib.import_module("joblib.numpy_pickle"), "")        # not shadowed in this process
        return _REAL
    # this file IS `joblib` here: import the real package with this directory off the path, park it under another name
    saved = list(sys.path)
    sys.modules.pop("joblib", None)
    try:
        sys.path = [p for p in sys.path if os.path.abspath(p or ".") != _HERE]
        importlib.import_module("joblib")
        importlib.import_module("joblib.numpy_pickle")
    finally:
        sys.path = saved
        for k in [k for k in sys.modules if k == "joblib" or k.startswith("joblib.")]:
            sys.modules["_real_" + k] = sys.modules.pop(k)
        if me is not None:
            sys.modules["joblib"] = me
    _REAL = (sys.modules["_real_joblib.numpy_pickle"], "_real_")
    return _REAL


class _State:
    def __init__(self, *args, **kwargs):
        self._args = args

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


class _Forest:
    def __init__(self, obj):
        self.classes_ = np.asarray(obj.classes_)
        self.trees = []
        for est in obj.estimators_:
            t = est.tree_
            nodes = t.nodes
            self.trees.append((np.asarray(nodes["left_child"]), np.asarray(nodes["right_child"]), np.asarray(nodes["feature"]),
                               np.asarray(nodes["threshold"], np.float64), np.asarray(t.values, np.float64)[:, 0, :]))

    def predict_proba(self, X):
        X = np.asarray(X, dtype=np.float32)
        if X.ndim == 1:
            X = X.reshape(1, -1)
        out = np.zeros((X.shape[0], len(self.classes_)), np.float64)
        for left, right, feat, thr, val in self.trees:
            for r in range(X.shape[0]):
                k = 0
                while left[k] != -1:
                    k = left[k] if X[r, feat[k]] <= thr[k] else right[k]
                v = val[k]
                out[r] += v / v.sum()
        return out / len(self.trees)

    def predict(self, X):
        return self.classes_.take(np.argmax(self.predict_proba(X), axis=1), axis=0)


def load(path):
    npk, prefix = _real_numpy_pickle()

    class Unpickler(npk.NumpyUnpickler):
        def find_class(self, module, name):
            if module.split(".")[0] == "sklearn":
                return type(name, (_State,), {})
            if module.split(".")[0] == "joblib":          # e.g. joblib.numpy_pickle.NumpyArrayWrapper: from the real package
                return getattr(sys.modules[prefix + module], name)
            return super().find_class(module, name)

    with open(path, "rb") as f:
        with npk._validate_fileobject_and_memmap(f, path, None) as (fobj, _):
            if isinstance(fobj, str):
                raise ValueError("old joblib pickle format is not supported")
            obj = Unpickler(path, fobj, ensure_native_byte_order=True, mmap_mode=None).load()
    return _Forest(obj)
