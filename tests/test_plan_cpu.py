"""Host logic of the tensor-core scan: the piece table (tc_ts_plan, scan_tc.cu) must cover every
(query block, tile) exactly once, give every piece its own candidate slot, and balance the SMs.
Needs no GPU: nmslib_b200_scan_plan is a pure host entry of the C ABI."""
import numpy as np
import pytest

import nmslib_zig_b200 as nb

CASES = [(10_000, 1_000_000, 10), (1_000, 10_000, 10), (1, 100, 5), (77, 3_001, 7), (256, 64, 3),
         (100_000, 1_000_000, 100), (38_000, 500_000, 10), (2_000, 100_000, 10), (5_000, 125_000, 10),
         (300, 20_000, 10), (12_544, 333_334, 10), (257, 65, 1), (10_000, 125_000, 10), (40_000, 10_000_000, 10)]


@pytest.mark.parametrize("nq,n,k", CASES)
@pytest.mark.parametrize("sms", [148, 132, 8])
def test_plan_covers_every_tile_once(nq, n, k, sms):
    pieces, n_cta, s_max = nb.scan_plan(nq, n, k, sms)
    blocks, tiles = (nq + 255) // 256, (n + 63) // 64
    cover = np.zeros((blocks, tiles), np.int32)
    per_cta = {}
    for cta, qb, t0, t1, slot in pieces:
        assert 0 <= cta < n_cta and 0 <= qb < blocks and 0 <= t0 < t1 <= tiles and 0 <= slot < s_max
        cover[qb, t0:t1] += 1
        per_cta.setdefault(cta, []).append((qb, slot))
    assert (cover == 1).all()
    assert len({(qb, slot) for _, qb, _, _, slot in pieces}) == len(pieces)   # one candidate list per piece
    assert max(len(v) for v in per_cta.values()) <= 8                          # TS_MAXP
    assert s_max <= 64                                                         # re-rank limit


def test_plan_balances_config2():
    """config 2 on 148 SMs: every SM busy, nobody more than 1 % above the mean."""
    pieces, n_cta, s_max = nb.scan_plan(10_000, 1_000_000, 10, 148)
    work = np.zeros(n_cta, np.int64)
    for cta, _, t0, t1, _ in pieces:
        work[cta] += t1 - t0
    assert n_cta == 148 and work.min() > 0
    assert work.max() <= 1.01 * work.mean()
    assert s_max <= 6


PAIR_CASES = [(10_000, 1_000_000, 100), (100_000, 1_250_000, 100), (20_000, 40_000, 100), (19_200, 30_000, 10),
              (64, 5_000, 10), (1, 300, 5), (300, 20_000, 233), (32_768, 1_000_000, 233), (5_000, 257, 10)]


@pytest.mark.parametrize("nq,n,k", PAIR_CASES)
@pytest.mark.parametrize("sms", [148, 132, 8])
def test_pair_plan_covers_every_tile_once(nq, n, k, sms):
    """Long rows run on CTA pairs over 256-row tiles (tc_scan_pair_kernel): same coverage rules, two candidate lists
    per piece, and the number of lists per query block within what the re-rank reads (64)."""
    pieces, n_pairs, lists = nb.scan_plan_pairs(nq, n, k, sms)
    blocks, tiles = (nq + 255) // 256, (n + 255) // 256
    cover = np.zeros((blocks, tiles), np.int32)
    per_pair = {}
    for pair, qb, t0, t1, slot in pieces:
        assert 0 <= pair < n_pairs and 0 <= qb < blocks and 0 <= t0 < t1 <= tiles and 0 <= 2 * slot + 1 < lists
        cover[qb, t0:t1] += 1
        per_pair.setdefault(pair, []).append((qb, slot))
    assert (cover == 1).all()
    assert len({(qb, slot) for _, qb, _, _, slot in pieces}) == len(pieces)
    assert max(len(v) for v in per_pair.values()) <= 8
    assert lists <= 64 and lists % 2 == 0


def test_pair_plan_balances_config5_shard():
    """config 5's per-GPU work (100 K queries x 1.25 M rows) on 74 pairs: whole waves of query blocks, then the
    tail blocks cut so that every pair of the last wave has work and nobody exceeds the mean of that wave by 2 %."""
    pieces, n_pairs, lists = nb.scan_plan_pairs(100_000, 1_250_000, 100, 148)
    work = np.zeros(n_pairs, np.int64)
    for pair, _, t0, t1, _ in pieces:
        work[pair] += t1 - t0
    tiles = (1_250_000 + 255) // 256
    whole = int(np.sum(work == tiles))
    assert whole == (391 // 74) * 74 and n_pairs == whole + 74 and work.min() > 0
    tail = work[whole:]
    assert tail.max() <= 1.02 * tail.mean()
