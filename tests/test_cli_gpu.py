"""`build/wepp detectPeaks` against the reference's own binary (oracle/_ref/wepp_ref: the reference's
translation units shim-compiled, see oracle/Makefile) on complete on-disk workspaces: same command line,
same input files, a deterministic stand-in for `freyja demix` on PATH (tests/wepp_dataset.py).
Every file the reference writes is compared — byte for byte where the reference's order is defined,
as sets of rows / sorted row members where it iterates a tbb::concurrent_hash_map or appends from
parallel threads (src/WEPP/arena.cpp:667-689, :814-824, :895-903)."""
import os
import shutil
import subprocess

import pytest

from tests import wepp_dataset as wd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "build", "wepp")
REF = os.path.join(ROOT, "oracle", "_ref", "wepp_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/wepp_ref not built"),
              pytest.mark.skipif(not os.path.exists(OURS), reason="build/wepp not built")]


def _run(binary, ws, root, **kw):
    env = wd.env_with_fake_freyja(dict(ws, bin=os.path.join(root, "bin")))
    env.update(kw.pop("env", {}))
    args = wd.cli_args(dict(ws, wepp_dir=os.path.join(root, "weppdir")), **kw)
    r = subprocess.run([binary] + args, cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (binary, r.stdout[-2000:], r.stderr[-2000:])
    return r


def _read(root, ws, kind, name):
    d = {"i": os.path.join(root, "intermediate", ws["dataset"]), "r": os.path.join(root, "results", ws["dataset"])}[kind]
    with open(os.path.join(d, name)) as f:
        return f.read()


def _rows_as_sets(text):
    out = {}
    for line in text.splitlines():
        parts = line.split(",")
        key = parts[0]
        out.setdefault(key, []).extend(parts[1:])
    return {k: sorted(v) for k, v in out.items()}


def compare_workspaces(a, b, ws):
    P = ws["prefix"]
    exact = [("i", f"{P}_checkpoint.txt"), ("i", f"{P}_barcodes.csv"), ("i", "freyja_output_latest.txt"),
             ("i", "residual_mutations.txt"), ("r", f"{P}_haplotype_abundance.csv"), ("r", f"{P}_haplotype_uncertainty.csv"),
             ("r", f"{P}_lineage_abundance.csv"), ("r", f"{P}_haplotypes.tsv"), ("r", f"{P}_sam_generation_called.txt")]
    for kind, name in exact:
        assert _read(a, ws, kind, name) == _read(b, ws, kind, name), name
    # the stand-in logs its command lines: same number of Freyja rounds, same arguments (modulo the root)
    la = _read(a, ws, "i", "freyja_calls.log").replace(a, "<root>")
    lb = _read(b, ws, "i", "freyja_calls.log").replace(b, "<root>")
    assert la == lb
    for name in (f"{P}_haplotype_reads.csv", f"{P}_mutation_reads.csv", f"{P}_mutation_haplotypes.csv"):
        ra, rb = _rows_as_sets(_read(a, ws, "r", name)), _rows_as_sets(_read(b, ws, "r", name))
        assert ra.keys() == rb.keys(), name
        for k in ra:
            assert ra[k] == rb[k], (name, k)
    ca = sorted(_read(a, ws, "r", f"{P}_haplotype_coverage.csv").splitlines())
    cb = sorted(_read(b, ws, "r", f"{P}_haplotype_coverage.csv").splitlines())
    assert ca == cb


CASES = {
    "small": dict(n_nodes=2500, genome=3000, n_reads=3000, seed=5),
    "nomask_plain_pb": dict(n_nodes=1500, genome=2000, n_reads=2000, seed=11, with_mask=False, tree_name="tree.pb"),
    "selective": dict(n_nodes=30000, genome=8000, n_reads=6000, seed=3, n_templates=12, n_amplicons=30),
}


@pytest.mark.parametrize("case", list(CASES))
def test_detect_peaks_matches_reference_binary(case, tmp_path):
    a, b = str(tmp_path / "ref"), str(tmp_path / "ours")
    ws = wd.make_workspace(a, **CASES[case])
    shutil.copytree(a, b)
    _run(REF, ws, a)
    # (one case also cuts the id sort of the load stage into runs, as an 8 M-node tree does by itself)
    _run(OURS, ws, b, env={"WEPP_SORT_RUN": "64"} if case == "nomask_plain_pb" else {})
    compare_workspaces(a, b, ws)


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0"])
def test_detect_peaks_read_sharded_matches_reference_binary(devices, tmp_path):
    """WEPP_DEVICES: the initial filter of `build/wepp detectPeaks` read-sharded over several ranks of one process
    (wepp_group; here the ranks share device 0 so that the single-GPU box covers the path — on a multi-GPU box
    WEPP_GPUS=N puts a rank on each device, tests/test_multigpu_peer.py).  Every output file as the reference's."""
    a, b = str(tmp_path / "ref"), str(tmp_path / "ours")
    ws = wd.make_workspace(a, **CASES["selective"])
    shutil.copytree(a, b)
    _run(REF, ws, a)
    r = _run(OURS, ws, b, env={"WEPP_DEVICES": devices})
    assert f"peak selection on {len(devices.split(','))} GPUs" in r.stdout
    compare_workspaces(a, b, ws)


def test_detect_peaks_without_lineages_and_few_survivors(tmp_path):
    """clade-idx -1 (the documented "no lineages" value, parsed as uint32 and wrapped) and a stand-in that keeps
    one haplotype in seven, so that the neighbour rounds of the post filter do real work."""
    a, b = str(tmp_path / "ref"), str(tmp_path / "ours")
    ws = wd.make_workspace(a, n_nodes=6000, genome=4000, n_reads=3000, seed=8, n_templates=25)
    shutil.copytree(a, b)
    env = {"FAKE_FREYJA_KEEP_MOD": "7", "FAKE_FREYJA_RESIDUALS": "30"}
    _run(REF, ws, a, clade_idx=-1, env=env)
    _run(OURS, ws, b, clade_idx=-1, env=env)
    compare_workspaces(a, b, ws)


def test_cli_exit_codes():
    """main.cpp:36-68: help and no command exit 0, an unknown command 1; a missing input is fatal (exit 1)."""
    assert subprocess.run([OURS, "help"], capture_output=True).returncode == 0
    assert subprocess.run([OURS], capture_output=True).returncode == 0
    assert subprocess.run([OURS, "frobnicate"], capture_output=True).returncode == 1
    assert subprocess.run([OURS, "detectPeaks", "--bogus", "1"], capture_output=True).returncode == 1
    r = subprocess.run([OURS, "detectPeaks", "-d", "nope", "-i", "none.pb", "-f", "none.fa", "-p", "x"], capture_output=True)
    assert r.returncode == 1
