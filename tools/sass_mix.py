#!/usr/bin/env python3
"""Static view of the library's kernels (no GPU needed): registers / shared / local memory per kernel from `cuobjdump
--dump-resource-usage`, and for every kernel the SASS instruction mix of its HOT LOOP -- the innermost loop (a backward branch whose
body holds no other backward branch) with the most instructions -- split into FP32, FP64, integer/logic, shuffle, shared, global/local,
control and other.  For the register-band sweeps the hot loop is one wavefront step of a lane = C cells, so `per cell` = loop length / C
is the static counterpart of ncu's executed thread-instructions per cell (profiles/r2_posterior_kernels_ncu_full.txt).
Usage: sass_mix.py [substring of the kernel name ...]   (default: the kernels of the bench step)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mlprobs_b200", "libmlprobs_b200.so")

CLASSES = [
    ("fp64", r"^(DADD|DMUL|DFMA|DSETP|DMNMX|D2F|F2D|D2I|I2D|DSEL|MUFU\.RCP64H|MUFU\.RSQ64H)"),
    ("fp32", r"^(FADD|FMUL|FFMA|FSETP|FMNMX|FSEL|FCHK|MUFU|F2I|I2F|F2F|FSET|F2FP|FRND)"),
    ("shfl", r"^(SHFL|VOTE|MATCH|REDUX|WARPSYNC)"),
    ("smem", r"^(LDS|STS|ATOMS|LDSM|UBLKCP|SYNCS)"),
    ("gmem", r"^(LDG|STG|LD\b|ST\b|LDL|STL|ATOMG|ATOM|RED|LDGSTS|LDGDEPBAR|DEPBAR|CCTL|MEMBAR|ERRBAR)"),
    ("ctrl", r"^(BRA|BSSY|BSYNC|EXIT|RET|CALL|BAR|BREAK|BMOV|NOP|YIELD|WARPSYNC|JMP|BRX|ENDCOLLECTIVE)"),
    ("int", r"^(IADD|IADD3|IMAD|IMNMX|ISETP|LOP|LOP3|SHF|SHL|SHR|LEA|SEL|MOV|PRMT|POPC|FLO|BREV|I2I|I2IP|IABS|PLOP3|P2R|R2P|S2R|S2UR|CS2R|ULDC|LDC|LDCU|U|R2UR|VIADD|VIMNMX|SGXT|BMSK|IDP|VABSDIFF)"),
]


def classify(op):
    for name, rx in CLASSES:
        if re.match(rx, op):
            return name
    return "other"


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def resources():
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function ([^:\s]+):\s*\n\s*(.*)", txt):
        res[m.group(1)] = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
    return res


def kernels():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
            body[cur].append((int(m.group(1), 16), ins))
    return body


def hot_loop(ins):
    """(start index, end index) of the innermost backward-branch loop with the most instructions, or None."""
    addr_to_idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, text) in enumerate(ins):
        m = re.match(r"BRA(?:\.\S+)*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?", text)
        t = re.search(r"0x([0-9a-f]+)", text) if text.startswith("BRA") else None
        if t:
            tgt = int(t.group(1), 16)
            if tgt <= a and tgt in addr_to_idx:
                loops.append((addr_to_idx[tgt], i))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    return max(inner, key=lambda l: l[1] - l[0]) if inner else None


def main():
    want = sys.argv[1:] or ["k_hmm_fwd_c", "k_hmm_bwd_c", "k_part_fwd_c", "k_part_rev_c", "k_part_fwd_s", "k_part_rev_s", "k_final_c", "k_loc_fwd_c", "k_loc_bwd_c",
                            "k_relax_blk", "k_transpose", "k_tree", "k_profile_posterior", "k_mea_wavefront"]
    res, body = resources(), kernels()
    names = demangle(list(body))
    print("%-44s %4s %6s %6s %6s | hot loop: %5s %8s | %5s %5s %5s %5s %5s %5s %5s %5s" %
          ("kernel", "regs", "shared", "local", "total", "instr", "per cell", "fp32", "fp64", "int", "shfl", "smem", "gmem", "ctrl", "other"))
    for k in sorted(body, key=lambda k: names[k]):
        nm = re.sub(r"\(anonymous namespace\)::|void |\(KArgs\)|\(.*\)$", "", names[k])
        if not any(w in nm for w in want):
            continue
        r = res.get(k, {})
        ins = body[k]
        hl = hot_loop(ins)
        if not hl:
            print("%-44s %4s %6s %6s %6d | no loop" % (nm[:44], r.get("REG", "?"), r.get("SHARED", "?"), r.get("LOCAL", "?"), len(ins)))
            continue
        loop = ins[hl[0]:hl[1] + 1]
        mix = collections.Counter(classify(t.split()[0]) for _, t in loop)
        c = re.search(r"<(\d+)", nm)
        per = "%.1f" % (len(loop) / int(c.group(1))) if c and "_c<" in nm or c and "_s<" in nm else "-"
        print("%-44s %4s %6s %6s %6d | %14d %8s | %5d %5d %5d %5d %5d %5d %5d %5d" %
              (nm[:44], r.get("REG", "?"), r.get("SHARED", "?"), r.get("LOCAL", "?"), len(ins), len(loop), per,
               mix["fp32"], mix["fp64"], mix["int"], mix["shfl"], mix["smem"], mix["gmem"], mix["ctrl"], mix["other"]))


if __name__ == "__main__":
    main()
