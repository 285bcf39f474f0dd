"""CPU: the semantic-dedup oracle against brute force (the reference's CuPy/cuML arithmetic cannot run here: parity with the
reference itself is unpinned, see oracle/dedup.py), and the host-side surface of the product module."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import dedup as od


def _brute(ids, emb, dist, eps):
    order = sorted(range(len(ids)), key=lambda i: (-np.float32(dist[i]), i))
    e = np.asarray(emb, np.float32)[order]
    e = e / np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-12).astype(np.float32)
    m = len(order)
    maxv, arg = np.full(m, -1.0, np.float32), np.full(m, -1, np.int64)
    for j in range(m):
        for i in range(j):
            s = np.float32(min(1.0, max(-1.0, float(np.dot(e[i], e[j])))))
            if s > maxv[j]:
                maxv[j], arg[j] = s, i
    if m:
        maxv[0], arg[0] = 0.0, 0
    arg = np.where(arg < 0, 0, arg)
    sid = np.asarray(ids)[order]
    return sid, sid[arg], maxv, int((maxv <= np.float32(1 - eps)).sum())


@pytest.mark.parametrize("m", [1, 2, 7, 130])
def test_pairwise_max_matches_brute_force(m):
    rng = np.random.default_rng(m)
    emb = rng.standard_normal((m, 32)).astype(np.float32)
    if m > 5:
        emb[5] = emb[1]  # exact duplicate
        emb[6] = 3.0 * emb[1]  # duplicate up to scale
    dist = rng.random(m).astype(np.float32)
    if m > 3:
        dist[3] = dist[2]  # tie in the sort key: stable order
    ids = np.array([f"clip-{i}" for i in range(m)])
    r = od.pairwise_max(ids, emb, dist, eps=0.05)
    sid, mid, mv, kept = _brute(ids, emb, dist, 0.05)
    assert list(r["id"]) == list(sid)
    np.testing.assert_allclose(r["cosine_sim_score"], mv, atol=2e-6)
    # where the chosen neighbour differs it must be an fp32 near-tie (matmul vs dot summation order), e.g. among duplicates
    e = r["sim_matrix_unit"]
    pos = {v: k for k, v in enumerate(r["id"])}
    for j in np.flatnonzero(r["max_id"] != mid):
        assert abs(float(e[pos[r["max_id"][j]]] @ e[j]) - float(e[pos[mid[j]]] @ e[j])) < 2e-6
    assert r["total"] == m and abs(r["kept"] - kept) <= 1


def test_pairwise_duplicates_are_pruned_and_first_index_wins():
    base = np.eye(16, dtype=np.float32)[:4]
    emb = np.concatenate([base, base[[2]], base[[2]]])  # rows 4 and 5 duplicate row 2
    dist = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.4], np.float32)  # already descending
    r = od.pairwise_max(np.arange(6), emb, dist, eps=0.01)
    assert list(r["max_id"][4:]) == [2, 2]  # not 4: the first of the equal maxima
    np.testing.assert_allclose(r["cosine_sim_score"][4:], 1.0)
    assert r["kept"] == 4 and r["total"] == 6
    assert r["cosine_sim_score"][0] == 0.0 and r["max_id"][0] == 0
    # orthogonal rows: similarity 0 > -1 -> index of the first earlier row
    assert list(r["max_id"][1:4]) == [0, 0, 0]


def test_assign_is_nearest_centroid():
    rng = np.random.default_rng(3)
    x = od.l2_normalize(rng.standard_normal((500, 24)))
    c = rng.standard_normal((9, 24)).astype(np.float32) * 0.7
    labels, cd = od.assign(x, c)
    d2 = ((x[:, None, :] - c[None]) ** 2).sum(-1)
    assert np.array_equal(labels, d2.argmin(1))
    cu = c / np.linalg.norm(c, axis=1, keepdims=True)
    np.testing.assert_allclose(cd, 1 - (x * cu[labels]).sum(1), atol=1e-6)


def test_product_module_imports_without_a_gpu():
    import cosmos_curate_b200.dedup as pd

    assert callable(pd.semdedup_cluster) and callable(pd.spherical_kmeans) and callable(pd.rowdot_argmax)


def _check_against_reference_golden(name, g, got_maxv, got_arg_pos):
    """got_* in the sorted order the reference works in; arg-max compared where the decision is not a float near-tie."""
    maxv, argi = g[name + "_maxv"], g[name + "_argi"]
    np.testing.assert_allclose(got_maxv, maxv, rtol=0, atol=3e-6)
    e = od.l2_normalize(g[name + "_emb"][np.argsort(-g[name + "_dist"], kind="stable")])
    argi0 = np.where(argi < 0, 0, argi)
    differ = np.flatnonzero(got_arg_pos != argi0)
    for j in differ:  # a different row may only be chosen if it is an equally good (within fp32 summation noise) earlier row
        assert got_arg_pos[j] < j and abs(float(e[got_arg_pos[j]] @ e[j]) - float(e[argi0[j]] @ e[j])) < 3e-6, j
    assert len(differ) <= max(1, len(maxv) // 200)


@pytest.mark.parametrize("name", ["multi_tile", "default_tile", "tiny"])
def test_oracle_pinned_to_reference_executed_dedup(name):
    """oracle/dedup.pairwise_max against the reference's OWN dedup array code (dedup_actor.py:404-466, run from its source with
    numpy standing in for cupy, oracle/ref_import.dedup_core -> tests/golden/dedup_ref.npz): start values, strict `>` across
    tiles, first arg-max inside a tile, the tril mask of the diagonal tile, clip, the legacy row-0 convention."""
    from conftest import load_golden

    g = load_golden("dedup_ref.npz")
    emb, dist = g[name + "_emb"], g[name + "_dist"]
    ids = np.arange(len(emb))
    r = od.pairwise_max(ids, emb, dist, eps=0.01)
    order = np.argsort(-dist, kind="stable")
    pos = {int(i): k for k, i in enumerate(order)}
    got_arg = np.array([pos[int(i)] for i in r["max_id"]])
    assert np.array_equal(r["id"], ids[order])
    _check_against_reference_golden(name, g, r["cosine_sim_score"], got_arg)
    thr = np.float32(0.99)
    safe = np.abs(g[name + "_maxv"] - thr) > 1e-5
    assert np.array_equal((r["cosine_sim_score"] <= thr)[safe], (g[name + "_maxv"] <= thr)[safe])
