"""quickprobs_b200 (mlprobs_b200/bin): the command-line drop-in for `quickprobs` over the C ABI."""
import os
import subprocess
import numpy as np
import pytest
from common import load_golden, split_seqs, HERE

CLI = os.path.join(os.path.dirname(HERE), "mlprobs_b200", "bin", "quickprobs_b200")
QP_EXE = CLI
CPNP_EXE = os.path.join(os.path.dirname(HERE), "mlprobs_b200", "bin", "c_p_np_aln_b200")


def fasta_text(headers, rows):
    out = []
    for h, r in zip(headers, rows):
        out.append(">" + h)
        out += [r[p:p + 60] for p in range(0, len(r), 60)]
    return "\n".join(out) + "\n"


def write_input(tmp_path, seqs):
    fa = tmp_path / "in.fa"
    heads = ["seq%d  some description " % i for i in range(len(seqs))]
    # lower case + Windows line ends + blank lines: all normalised by the loader (SequenceIO.cpp:70-155)
    fa.write_bytes(b"".join(b">" + h.encode() + b"\r\n\n" + s.lower() + b"\r\n" for h, s in zip(heads, seqs)))
    return str(fa), [h.strip() for h in heads]


def test_cli_is_built():
    assert os.path.exists(CLI), "run __graft_entry__.build()"


def test_cli_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa, _ = write_input(tmp_path, [b"ACDEFGHIK", b"ACDEFGHIK"])
    r = subprocess.run([CLI, fa], capture_output=True, text=True)
    assert r.returncode != 0 and "CUDA device is required" in r.stderr and r.stdout == ""


def test_cli_rejects_illegal_characters_and_unknown_options(tmp_path):
    fa = tmp_path / "bad.fa"
    fa.write_text(">a\nACD*EF\n>b\nACDEF\n")
    r = subprocess.run([CLI, str(fa)], capture_output=True, text=True)
    assert r.returncode == 255 and "illegal sequence character:*" in r.stdout and "Illegal characters" in r.stderr
    r = subprocess.run([CLI, str(fa), "--nucleotide"], capture_output=True, text=True)
    assert r.returncode == 2 and "unsupported option" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["qp_sup139", "qp_676s4"])
def test_cli_reproduces_the_reference_alignment(tmp_path, name):
    d = load_golden(name)
    seqs = split_seqs(d)
    fa, heads = write_input(tmp_path, seqs)
    want = fasta_text(heads, [r.tobytes().decode() for r in d["msa"]])
    out = tmp_path / "out.fa"
    r = subprocess.run([CLI, fa, "-o", str(out), "-t", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == want
    r = subprocess.run([CLI, fa], capture_output=True, text=True)      # stdout when no -o is given
    assert r.returncode == 0 and r.stdout == want


@pytest.mark.gpu
def test_cli_directory_mode_aligns_every_file_with_one_context(tmp_path):
    """QuickProbs' bulk entry (Configuration.cpp:248-267): input directory -> output directory, same file names."""
    indir = tmp_path / "in"; outdir = tmp_path / "out"
    indir.mkdir(); outdir.mkdir()
    want = {}
    for name in ("qp_sup139", "qp_sup002"):
        d = load_golden(name)
        fa, heads = write_input(tmp_path, split_seqs(d))
        os.rename(fa, indir / (name + ".fa"))
        want[name + ".fa"] = fasta_text(heads, [r.tobytes().decode() for r in d["msa"]])
    r = subprocess.run([CLI, str(indir), "-o", str(outdir)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for nm, text in want.items():
        assert (outdir / nm).read_text() == text


@pytest.mark.gpu
def test_cli_single_sequence_is_echoed(tmp_path):
    fa, heads = write_input(tmp_path, [b"ACDEFGHIKLMNPQRSTVWY" * 4])
    r = subprocess.run([CLI, fa], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == fasta_text(heads, [("ACDEFGHIKLMNPQRSTVWY" * 4)])


CPNP = os.path.join(os.path.dirname(HERE), "mlprobs_b200", "bin", "c_p_np_aln_b200")


def test_cpnp_cli_is_built_and_refuses_without_gpu(tmp_path):
    assert os.path.exists(CPNP), "run __graft_entry__.build()"
    fa, _ = write_input(tmp_path, [b"ACDEFGHIK", b"ACDEFGHIK"])
    r = subprocess.run([CPNP, "-p", "2", fa], capture_output=True, text=True)
    assert r.returncode == 1 and "integer must be 0 or 1" in r.stderr and r.stdout == ""
    import torch
    if not torch.cuda.is_available():
        for prog in ("0", "1"):
            r = subprocess.run([CPNP, "-p", prog, fa], capture_output=True, text=True)
            assert r.returncode != 0 and "CUDA device is required" in r.stderr and r.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cpnp_p1_BB12003", "cpnp_p1_676s4"])
def test_cpnp_cli_non_progressive_reproduces_the_reference(tmp_path, name):
    """`c_p_np_aln -p 1 file`; --seed stands in for the wall clock the reference seeds its refinement sweeps with."""
    d = load_golden(name)
    seqs = split_seqs(d)
    fa, heads = write_input(tmp_path, seqs)
    want = fasta_text(heads, [r.tobytes().decode() for r in d["msa"]])
    r = subprocess.run([CPNP, "-p", "1", "--seed", str(int(d["fixtime"][0])), fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want and r.stderr == ""
    want0 = fasta_text(heads, [r_.tobytes().decode() for r_ in d["msa_ir0"]])
    r = subprocess.run([CPNP, "-p", "1", "-ir", "0", fa], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == want0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cpnp_sup002_ref", "cpnp_676s4_ref"])
def test_cpnp_cli_reproduces_the_reference(tmp_path, name):
    d = load_golden(name)
    seqs = split_seqs(d)
    fa, heads = write_input(tmp_path, seqs)
    want = fasta_text([heads[i] for i in d["msa_order"]], [r.tobytes().decode() for r in d["msa"]])
    r = subprocess.run([CPNP, "-p", "0", fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want
    if name == "cpnp_sup002_ref":           # the -G feature line (20 standard letters only)
        r = subprocess.run([CPNP, "-G", fa], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout == d["gline"].tobytes().decode() + "\n"


def test_cpnp_cli_input_errors_match_the_reference_messages(tmp_path):
    """Sequence::Sequence / MSA::ParseParams error behaviour (Sequence.h:96-112, MSA.cpp:248-435): exit status 1,
    `ERROR:` on stderr, nothing on stdout -- MLProbs treats the non-zero status as failure.  These paths end before any CUDA
    call, so they are checked without a GPU too."""
    missing = str(tmp_path / "nope.fa")
    r = subprocess.run([CPNP, "-p", "0", missing], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout == "" and r.stderr == "ERROR: Could not open file '%s' for reading.\n" % missing
    bad = tmp_path / "bad.fa"
    bad.write_text(">a\nACD1EF\n>b\nACDEF\n")
    r = subprocess.run([CPNP, "-p", "1", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout == "" and r.stderr == "ERROR: Unknown character encountered: 1\n"
    empty = tmp_path / "empty.fa"
    empty.write_text("\n\n")
    r = subprocess.run([CPNP, "-G", str(empty)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout == "" and r.stderr == "ERROR: No sequences read.\n"
    r = subprocess.run([CPNP, "-p"], capture_output=True, text=True)
    assert r.returncode == 1 and "Must specify a value after option" in r.stderr
    r = subprocess.run([CPNP], capture_output=True, text=True)
    assert r.returncode != 0 and r.stdout == ""
    single = tmp_path / "single.fa"
    single.write_text(">a\nACDEFGHIK\n")
    for mode in (["-p", "0"], ["-p", "1"], ["-G"]):      # the reference crashes here (exit 139 / 136, empty stdout); MLProbs falls back
        r = subprocess.run([CPNP] + mode + [str(single)], capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout == "" and "at least two sequences" in r.stderr
    # option handling as MSA::ParseParams: unknown option and -version end with status 1, the options the reference parses
    # without any effect on its output are accepted
    ok = tmp_path / "ok.fa"
    ok.write_text(">a\nACDEF\n>b\nACDEF\n")
    r = subprocess.run([CPNP, "--bogus", str(ok)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout == "" and r.stderr == "ERROR: Unrecognized option: --bogus\n"
    r = subprocess.run([CPNP, "-version"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout == "" and r.stderr != ""
    import torch
    if not torch.cuda.is_available():       # accepted options and several input files get as far as the device check
        r = subprocess.run([CPNP, "-clustalw", "-timeon", "-p", "0", str(ok), str(ok)], capture_output=True, text=True)
        assert r.returncode != 0 and "CUDA device is required" in r.stderr and r.stdout == ""


def test_quickprobs_cli_option_syntax(tmp_path):
    """ProgramOptions::parse (Common/ProgramOptions.cpp:15-48): options take any number of leading dashes and a long or short
    name; without the positional argument the reference returns 0 and writes nothing to stdout (Console/main.cpp:33-37)."""
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == ""
    missing = str(tmp_path / "nope.fa")
    for opt in ("-o", "--o", "---outfile", "-outfile"):
        r = subprocess.run([CLI, opt, str(tmp_path / "out.fa"), "-t", "3", "--platform", "0", "--mem-limit", "100", missing], capture_output=True, text=True)
        assert r.returncode == 255 and "unable to open input file" in r.stderr and r.stdout == "", opt
    r = subprocess.run([CLI, "-l", missing], capture_output=True, text=True)
    assert r.returncode == 2 and "unsupported option" in r.stderr


@pytest.mark.gpu
def test_persistent_process_mode_gives_the_stand_alone_output(tmp_path):
    """MLP_B200_SERVER=1 (csrc/serve.h): the executable hands its command line to a server process of the same program that
    keeps one CUDA context; outputs, exit status and error messages are those of the stand-alone run."""
    from mlprobs_b200 import synth
    seqs = synth.family(7, 60, seed=31)
    fa = tmp_path / "in.fa"
    fa.write_text("".join(">s%d\n%s\n" % (i, s.decode()) for i, s in enumerate(seqs)))
    env = dict(os.environ, MLP_B200_SERVER="1", MLP_B200_SERVER_IDLE="3")
    for exe, args in ((QP_EXE, []), (CPNP_EXE, ["-p", "0"]), (CPNP_EXE, ["-G"]), (CPNP_EXE, ["-p", "1", "--seed", "5"])):
        alone = subprocess.run([exe] + args + [str(fa)], capture_output=True)
        for _ in range(2):                                  # first call starts the server, second re-uses it
            served = subprocess.run([exe] + args + [str(fa)], capture_output=True, env=env)
            assert served.returncode == alone.returncode == 0
            assert served.stdout == alone.stdout and len(alone.stdout) > 0
    # errors travel too: unknown option and a missing file keep their message and status
    for exe, args in ((CPNP_EXE, ["-zzz", str(fa)]), (QP_EXE, [str(tmp_path / "missing.fa")])):
        alone = subprocess.run([exe] + args, capture_output=True)
        served = subprocess.run([exe] + args, capture_output=True, env=env)
        assert served.returncode == alone.returncode != 0 and served.stderr == alone.stderr


def test_persistent_process_mode_protocol_without_a_gpu(tmp_path):
    """The client/server plumbing of csrc/serve.h needs no device to be exercised: without a GPU the server's run ends in the same
    'a CUDA device is required' error, input errors keep their message and exit status, and the server goes away when idle."""
    import time
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    fa, _ = write_input(tmp_path, [b"ACDEFGHIK", b"ACDEFGHIK", b"ACDEFGHIKL"])
    env = dict(os.environ, MLP_B200_SERVER="1", MLP_B200_SERVER_IDLE="2")
    sock = "/tmp/mlprobs_b200_%d_%s.sock" % (os.getuid(), "quickprobs_b200")
    for exe, args in ((QP_EXE, [fa]), (CPNP_EXE, ["-p", "0", fa]), (CPNP_EXE, ["-zzz", fa]), (QP_EXE, [str(tmp_path / "missing.fa")])):
        alone = subprocess.run([exe] + args, capture_output=True, text=True)
        for _ in range(2):
            served = subprocess.run([exe] + args, capture_output=True, text=True, env=env)
            assert served.returncode == alone.returncode != 0
            assert served.stderr == alone.stderr and served.stdout == alone.stdout
    assert os.path.exists(sock)                      # a server was started and is listening
    time.sleep(4)
    assert not os.path.exists(sock)                  # idle for MLP_B200_SERVER_IDLE seconds: it removed its socket and left
