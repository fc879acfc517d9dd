"""Host-side launch planners (no GPU): which GEMM kernel / tiling and which attention work items a problem gets.
The planners are pure host logic behind the C ABI (ltx2_gemm_plan, ltx2_attention_plan); with no device visible they
assume the B200's 148 SMs."""
import ctypes as C
import os
import random

import pytest

STANDARD, TRANSPOSED, PAIR = 0, 1, 2
EPI_BF16, EPI_GELU, EPI_F32, EPI_RESIDUAL = 0, 1, 2, 3


def gemm_plan(M, N, K, mode=EPI_BF16, max_splits=1):
    from ltx2_b200._lib import check, lib
    o = (C.c_int32 * 6)()
    check(lib().ltx2_gemm_plan(M, N, K, mode, max_splits, o), "ltx2_gemm_plan")
    return dict(kernel=o[0], bn=o[1], splits=o[2], tile_w=o[3], last_w=o[4], num_t=o[5])


def attention_plan(Tq, BH):
    from ltx2_b200._lib import check, lib
    pairs, ctas = C.c_int32(), C.c_int32()
    check(lib().ltx2_attention_plan(Tq, BH, C.byref(pairs), C.byref(ctas)), "ltx2_attention_plan")
    return pairs.value, ctas.value


@pytest.fixture(autouse=True)
def _no_overrides(monkeypatch):
    for k in ("LTX2_GEMM_T", "LTX2_GEMM_2CTA", "LTX2_ATTN_PAIRS"):
        monkeypatch.delenv(k, raising=False)


def test_full_size_problems_stay_on_the_standard_kernel():
    """The single-GPU bench shapes (3456 token rows): every linear of a block runs as 128 x 256 tiles, unsplit."""
    for N, K, mode in [(12288, 4096, EPI_BF16), (4096, 4096, EPI_RESIDUAL), (4096, 4096, EPI_BF16),
                       (16384, 4096, EPI_GELU), (4096, 16384, EPI_RESIDUAL)]:
        p = gemm_plan(3456, N, K, mode, 8 if mode == EPI_RESIDUAL else 1)
        assert (p["kernel"], p["bn"], p["splits"]) == (STANDARD, 256, 1), (N, K, mode, p)


def test_context_parallel_shards_get_shard_shaped_tiles():
    """432 / 864 / 1728 token rows: the wide bf16 projections move to SM-pair tiles that cover the tokens without a
    wasted 128-row tile; the residual GEMMs stay on the standard kernel and split K instead."""
    for M in (432, 864, 1728):
        for N, mode in ((12288, EPI_BF16), (16384, EPI_GELU)):
            p = gemm_plan(M, N, 4096, mode)
            assert p["kernel"] == PAIR, (M, N, p)
            assert p["tile_w"] % 32 == 0 and p["last_w"] % 32 == 0 and 0 < p["last_w"] <= p["tile_w"] <= 256
            covered = p["tile_w"] * (p["num_t"] - 1) + p["last_w"]
            assert M <= covered < M + 32, (M, p)
    p = gemm_plan(432, 4096, 4096, EPI_RESIDUAL, 8)
    assert p["kernel"] == STANDARD and 2 <= p["splits"] <= 8
    p = gemm_plan(432, 4096, 16384, EPI_RESIDUAL, 8)
    assert p["kernel"] == STANDARD and 2 <= p["splits"] <= 8
    assert gemm_plan(432, 4096, 4096, EPI_RESIDUAL, 1)["splits"] == 1          # split-K is opt-in per call


def test_forced_kernels_and_invariants(monkeypatch):
    rng = random.Random(0)
    for force_env, kernel, step in (("LTX2_GEMM_T", TRANSPOSED, 16), ("LTX2_GEMM_2CTA", PAIR, 32)):
        monkeypatch.setenv(force_env, "2")
        for _ in range(200):
            M = rng.randint(64, 5000)
            N = 256 * rng.randint(1, 80)
            K = 64 * rng.randint(2, 300)
            p = gemm_plan(M, N, K, rng.choice([EPI_BF16, EPI_GELU, EPI_F32]))
            assert p["kernel"] == kernel, (M, N, K, p)
            assert p["tile_w"] % step == 0 and 0 < p["tile_w"] <= 256 and p["num_t"] >= 1
            assert p["tile_w"] * (p["num_t"] - 1) < M <= p["tile_w"] * p["num_t"], (M, p)   # every tile has tokens
        monkeypatch.delenv(force_env)
    monkeypatch.setenv("LTX2_GEMM_T", "0")
    monkeypatch.setenv("LTX2_GEMM_2CTA", "0")
    for M, N in ((432, 12288), (100, 512), (1728, 16384)):
        assert gemm_plan(M, N, 4096)["kernel"] == STANDARD
    assert gemm_plan(100, 96, 640)["bn"] == 32 and gemm_plan(100, 192, 640)["bn"] == 64


def test_attention_work_items():
    """3456 queries x 32 heads: 13 pair items + 1 split-KV item per head = 448 CTAs, three waves of pairs on 148 SMs.
    Small grids (the 4 heads of an 8-way context-parallel rank) run as split-KV items only."""
    assert attention_plan(3456, 32) == (13, 448)
    assert attention_plan(3456, 4) == (0, 108)
    assert attention_plan(128, 2) == (0, 2)
    rng = random.Random(1)
    for _ in range(300):
        Tq, BH = rng.randint(1, 20000), rng.randint(1, 128)
        pairs, ctas = attention_plan(Tq, BH)
        n_q = (Tq + 127) // 128
        assert 0 <= pairs <= n_q // 2 and ctas == BH * (n_q - pairs)


def sm_pair_plan(Tq, Tk, BH):
    from ltx2_b200._lib import check, lib
    c, s = C.c_int32(), C.c_int32()
    check(lib().ltx2_attention_sm_pair_plan(Tq, Tk, BH, C.byref(c), C.byref(s)), "ltx2_attention_sm_pair_plan")
    return c.value, s.value


def test_sm_pair_attention_schedule():
    """The persistent SM-pair attention kernel: 32 heads x 14 query-tile pairs = 448 items do not fit 74 SM pairs in six
    rounds, so the (item, key block) space is cut into 74 equal ranges (stream-K); a rank of 8 (4 heads, 56 items) spreads
    its 1512 key blocks over all 74 pairs; few short items stay whole."""
    os.environ.pop("LTX2_ATTN_SPLIT", None)
    assert sm_pair_plan(3456, 3456, 32) == (74, 1)
    assert sm_pair_plan(3456, 3456, 4) == (74, 1)
    assert sm_pair_plan(432, 1024, 32) == (64, 0)          # 64 items of 8 blocks: one each
    assert sm_pair_plan(256, 128, 2) == (2, 0)
    rng = random.Random(2)
    for _ in range(300):
        Tq, Tk, BH = rng.randint(129, 20000), rng.randint(1, 20000), rng.randint(1, 64)
        clusters, split = sm_pair_plan(Tq, Tk, BH)
        n_items = BH * (((Tq + 127) // 128 + 1) // 2)
        nblk = (Tk + 127) // 128
        assert 1 <= clusters <= 74
        if split:
            # a range is at least four key blocks and a third of an item: at most four parts per item
            assert n_items * nblk // clusters >= max(4, nblk // 3)
        else:
            assert clusters <= n_items


def sm_pair_segments(Tq, Tk, BH, cluster):
    from ltx2_b200._lib import lib
    buf = (C.c_int32 * (7 * 64))()
    n = lib().ltx2_attention_sm_pair_segments(Tq, Tk, BH, cluster, buf, 64)
    assert 0 <= n <= 64
    return [tuple(buf[7 * i:7 * i + 7]) for i in range(n)]


@pytest.mark.parametrize("force_split", [None, "1"])
def test_sm_pair_attention_segments_cover_every_key_block_once(monkeypatch, force_split):
    """The device-side walk of the stream-K decomposition (shared host/device code): over all clusters every
    (item, key block) is visited exactly once, the parts of a split item are numbered 0..parts-1 in key order, every
    part knows the scratch slot of part 0, and no two partial segments share a scratch slot."""
    if force_split is None:
        monkeypatch.delenv("LTX2_ATTN_SPLIT", raising=False)
    else:
        monkeypatch.setenv("LTX2_ATTN_SPLIT", force_split)
    rng = random.Random(3)
    shapes = [(3456, 3456, 32), (3456, 3456, 4), (3456, 1024, 32), (12288, 12288, 8), (300, 200, 3), (384, 1000, 2)]
    shapes += [(rng.randint(129, 6000), rng.randint(1, 6000), rng.randint(1, 40)) for _ in range(25)]
    for Tq, Tk, BH in shapes:
        clusters, split = sm_pair_plan(Tq, Tk, BH)
        n_items = BH * (((Tq + 127) // 128 + 1) // 2)
        nblk = (Tk + 127) // 128
        seen = {}
        slots = set()
        parts_of = {}
        for c in range(clusters):
            for item, kb0, kb1, part, parts, slot, slot0 in sm_pair_segments(Tq, Tk, BH, c):
                assert 0 <= item < n_items and 0 <= kb0 < kb1 <= nblk
                for kb in range(kb0, kb1):
                    assert (item, kb) not in seen, (Tq, Tk, BH, item, kb)
                    seen[(item, kb)] = c
                if parts == 1:
                    assert (kb0, kb1) == (0, nblk) and slot == -1
                else:
                    assert split and 0 <= part < parts <= 8
                    assert slot not in slots
                    slots.add(slot)
                    parts_of.setdefault(item, []).append((part, kb0, kb1, slot, slot0, c))
        assert len(seen) == n_items * nblk, (Tq, Tk, BH)
        for item, ps in parts_of.items():
            ps.sort()
            assert [p[0] for p in ps] == list(range(len(ps))) and len(ps) == max(2, len(ps))
            assert ps[0][1] == 0 and ps[-1][2] == nblk and all(a[2] == b[1] for a, b in zip(ps, ps[1:]))
            assert all(p[4] == ps[0][3] for p in ps)                       # everyone knows part 0's slot
            assert all(p[3] == 2 * p[5] for p in ps[1:])                   # parts >= 1 start their cluster's range
        assert lib_rejects_cluster(Tq, Tk, BH, clusters)


def lib_rejects_cluster(Tq, Tk, BH, cluster):
    from ltx2_b200._lib import lib
    buf = (C.c_int32 * 7)()
    return lib().ltx2_attention_sm_pair_segments(Tq, Tk, BH, cluster, buf, 1) < 0


def test_planners_reject_bad_arguments():
    from ltx2_b200._lib import lib
    o = (C.c_int32 * 6)()
    assert lib().ltx2_gemm_plan(0, 256, 64, 0, 1, o) != 0
    assert lib().ltx2_attention_plan(128, 0, C.byref(C.c_int32()), None) != 0
