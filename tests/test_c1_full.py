"""The FULL C1 configuration (50,000 nodes x 200,000 reads, RSV-A primer scheme) against the reference's own object
code: tests/golden/c1_full.npz was written by tests/golden/make_c1.py from oracle/_ref/libwepp_ref.so
(initial_filter.cpp:139-239).  All integers exact — min parsimony, multiplicity, every cell of
mapped_read_counts[N][50] (SHA-256 of the matrix) — score and dist_divergence to 1e-9 relative (SCORE_EPSILON,
config.hpp:15).  Every GPU placement path is held to the same fixture."""
import hashlib
import os

import numpy as np
import pytest

from tests.golden import make_c1

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_full.npz")


@pytest.fixture(scope="module")
def c1():
    arena, reads = make_c1.inputs()
    g = np.load(GOLD)
    assert str(g["digest"]) == make_c1.digest(arena, reads), "the seeded C1 inputs changed: regenerate tests/golden/c1_full.npz"
    return arena, reads, g


def test_c1_fixture_shape(c1):
    arena, reads, g = c1
    assert arena.n_nodes == 50_000 and reads.n_reads == 200_000 and arena.genome_size == 15222
    assert g["max_parsimony"].shape == (200_000,) and g["score"].shape == (50_000,)
    assert g["multiplicity"].min() >= 1


def test_c1_oracle_sample_matches_reference(c1):
    """The CPU restatement (with the reference's range trees) on a 4,000-read slice of C1: the oracle is pinned on the
    reference at this shape too."""
    import oracle
    arena, reads, g = c1
    sub = reads.slice(100_000, 104_000)
    o = oracle.cartesian_map(arena, sub, None, n_threads=os.cpu_count() or 1, want_node=False, range_trees=True, range_reads=reads)
    assert np.array_equal(o["max_parsimony"], g["max_parsimony"][100_000:104_000])
    assert np.array_equal(o["multiplicity"], g["multiplicity"][100_000:104_000])


@pytest.mark.gpu
@pytest.mark.parametrize("env,path", [({}, 2), ({"WEPP_DELTA_PLACE": "0"}, 1), ({"WEPP_STATE_PLACE": "0"}, 0),
                                      ({"WEPP_NODE_TILES": "0"}, 2)])
def test_c1_full_on_gpu_matches_reference(c1, env, path, monkeypatch):
    from wepp_b200.placement import Placer
    arena, reads, g = c1
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    p = Placer(0)
    p.set_arena(arena)
    p.set_reads(reads)
    p.place(0, 0)
    assert p.stats()["place_path"] == path
    mp, mu = p.read_results()
    sc, ct = p.node_results()
    sc2, dv = p.node_summary()
    p.close()
    assert np.array_equal(mp, g["max_parsimony"].astype(np.int32))
    assert np.array_equal(mu, g["multiplicity"])
    ct = np.ascontiguousarray(ct, dtype=np.int32)
    assert np.array_equal(ct.sum(axis=0, dtype=np.int64), g["counts_colsum"])
    assert np.array_equal(ct[::16], g["counts_rows_every_16"])
    assert hashlib.sha256(ct.tobytes()).hexdigest() == str(g["counts_sha256"])
    np.testing.assert_allclose(sc, g["score"], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(sc2, g["score"], rtol=1e-9, atol=1e-15)
    # dist_divergence (initial_filter.cpp:214-231) = bins with counts / true_read_counts > 0.5 % over bins with reads.
    # The reference session's true_read_counts also hold the all-reference cover reads that keep every site covered
    # (make_c1.cover_reads; they are not placed), the GPU's only the placed reads: the restated formula must give the
    # reference's values with the former and the GPU's with the latter.
    def divergence(true_counts):
        with np.errstate(divide="ignore", invalid="ignore"):
            over = (ct.astype(np.float64) / true_counts.astype(np.float64)[None, :]) > 0.5 / 100
        return over.sum(axis=1) / float((true_counts != 0).sum())
    bin_size = arena.genome_size // 50
    cover = make_c1.cover_reads(arena, reads)
    t_placed = np.bincount(np.minimum(reads.start // bin_size, 49), weights=reads.degree, minlength=50).astype(np.int64)
    t_all = np.bincount(np.minimum(cover.start // bin_size, 49), weights=cover.degree, minlength=50).astype(np.int64)
    np.testing.assert_allclose(divergence(t_all), g["dist_divergence"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(dv, divergence(t_placed), rtol=1e-12, atol=0)
