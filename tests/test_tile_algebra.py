"""CPU model of the carry/count algebra the CUDA sweep kernel uses (blt_b200/csrc/kernels.cu),
checked against the reference's sequential sweep (oracle/py_model.py).

The kernel cannot run without a GPU, so this file re-derives, in plain Python with the kernel's own
variable names, every non-obvious step: segment start bits, identity/constant carry functions,
the "first non-identity segment" delta, the tile aggregate, and the windowed decoupled look-back
fold.  Small segment/tile/window sizes make identity segments, identity tiles and multi-window
look-backs common in random tests.
"""
import random

import pytest

from oracle import py_model as pm


def start_bits(m, cin, seg):
    """kernels.cu start_bits(): run-parity trick."""
    mask = (1 << 32) - 1
    mm = m & ~cin & mask
    s = mm & ~(mm << 1) & mask
    e = mm & ~((mm + (s & 0x55555555)) & mask) & mask
    return ((e & 0x55555555) | (mm & ~e & 0xAAAAAAAA)) & ((1 << seg) - 1)


def seg_fn(m, seg):
    """(identity, const carry_out) of one segment."""
    all_ones = (1 << seg) - 1
    if m == all_ones:
        return True, 0
    lead = 0
    for j in range(seg - 1, -1, -1):
        if (m >> j) & 1:
            lead += 1
        else:
            break
    return False, lead & 1


def seg_count(m, vm, cin, seg):
    st = start_bits(m, cin, seg)
    em = vm & ~((st << 1) | cin) & ((1 << seg) - 1)
    return bin(em).count("1"), st, em


class Tile:
    """What one group computes before the look-back (tile carry_in assumed 0)."""

    def __init__(self, ms, vms, seg):
        self.ms, self.vms, self.seg = ms, vms, seg
        self.cin0, self.dep, self.cnt0 = [], [], []
        carry, dep = 0, True
        self.f_idx, self.f_delta = None, 0
        for i, (m, vm) in enumerate(zip(ms, vms)):
            ident, const = seg_fn(m, seg)
            self.cin0.append(carry)
            self.dep.append(dep)
            c, _, _ = seg_count(m, vm, carry, seg)
            self.cnt0.append(c)
            if dep and not ident:
                c1, _, _ = seg_count(m, vm, 1, seg)
                self.f_idx, self.f_delta = i, c - c1
            if not ident:
                carry, dep = const, False
        self.tile_id = all(seg_fn(m, seg)[0] for m in ms)
        self.tile_const = 0 if self.tile_id else carry
        self.total0 = sum(self.cnt0)
        assert self.f_delta in (0, 1)

    def aggregate(self):
        return dict(state="A", id=self.tile_id, const=self.tile_const, delta=self.f_delta, cnt0=self.total0)

    def emit(self, tile_cin, tokens_in, values):
        """Phase C: final carry per segment, positions, tokens."""
        out = []
        for i, (m, vm) in enumerate(zip(self.ms, self.vms)):
            cin = tile_cin if self.dep[i] else self.cin0[i]
            cnt, st, em = seg_count(m, vm, cin, self.seg)
            pos = sum(self.cnt0[:i]) - (self.f_delta if (tile_cin and self.f_idx is not None and i > self.f_idx) else 0)
            assert pos == len(out), (pos, len(out))
            for j in range(self.seg):
                if (em >> j) & 1:
                    k = i * self.seg + j
                    out.append(values[k] if (st >> j) & 1 else tokens_in[k])
        total = self.total0 - (self.f_delta if tile_cin else 0)
        assert total == len(out)
        return out


def lookback(status, tile, W):
    """kernels.cu decoupled_lookback() with a W-lane window, all descriptors already published."""
    run_id, run_const, run_delta, run_cnt0 = True, 0, 0, 0
    j = tile - 1
    while True:
        lanes = []
        for lane in range(W):
            idx = j - lane
            lanes.append(status[idx] if idx >= 0 else dict(state="P", const=0, count=0))
        p = next((l for l, s in enumerate(lanes) if s["state"] == "P"), W)
        act = range(0, min(p, W - 1) + 1)
        nonid = [l for l in act if not (l < p and lanes[l]["id"])]
        wsum = 0
        for lane in range(W):
            if lane < p:
                above = [l for l in nonid if l > lane]
                cin = lanes[min(above)]["const"] if above else 0
                wsum += lanes[lane]["cnt0"] - (1 if (cin and lanes[lane]["delta"]) else 0)
        w_id = not nonid
        w_const = 0 if w_id else lanes[min(nonid)]["const"]
        w_delta = 0
        if p == W and not w_id:
            w_delta = lanes[max(nonid)]["delta"]
        c_mid0 = 0 if w_id else w_const
        new_cnt0 = wsum + run_cnt0 - (1 if (c_mid0 and run_delta) else 0)
        new_delta = run_delta if w_id else w_delta
        new_const = w_const if run_id else run_const
        new_id = w_id and run_id
        if p < W:
            return new_const, lanes[p]["count"] + new_cnt0
        run_id, run_const, run_delta, run_cnt0 = new_id, new_const, new_delta, new_cnt0
        j -= W


def run_tiled(tokens, merges, chunk, seg, segs_per_tile, W, publish_prefix_every=1):
    n = len(tokens)
    m_all = [0] * n
    for i in range(n - 1):
        if (i + 1) % chunk != 0 and (tokens[i], tokens[i + 1]) in merges:
            m_all[i] = 1
    values = [merges.get((tokens[i], tokens[i + 1]), None) if i + 1 < n else None for i in range(n)]
    tile_elems = seg * segs_per_tile
    n_tiles = (n + tile_elems - 1) // tile_elems
    status, out = [], []
    for t in range(n_tiles):
        ms, vms = [], []
        for s in range(segs_per_tile):
            g = t * tile_elems + s * seg
            m = sum(m_all[g + j] << j for j in range(seg) if g + j < n)
            vm = sum(1 << j for j in range(seg) if g + j < n)
            ms.append(m)
            vms.append(vm)
        tile = Tile(ms, vms, seg)
        if t == 0:
            cin, base = 0, 0
        else:
            cin, base = lookback(status, t, W)
        assert base == len(out), (t, base, len(out))
        toks = tile.emit(cin, tokens[t * tile_elems:] + [0] * tile_elems, values[t * tile_elems:] + [None] * tile_elems)
        out += toks
        c_out = cin if tile.tile_id else tile.tile_const
        # Mimic in-flight tiles: only some predecessors have turned their AGGREGATE into a PREFIX.
        if t % publish_prefix_every == 0:
            status.append(dict(state="P", const=c_out, count=base + len(toks)))
        else:
            status.append(tile.aggregate())
    return out


@pytest.mark.parametrize("seg,segs_per_tile,W", [(4, 2, 2), (4, 4, 3), (16, 2, 4), (8, 3, 32), (2, 2, 2)])
def test_tiled_sweep_equals_sequential(seg, segs_per_tile, W):
    rng = random.Random(seg * 1000 + segs_per_tile * 10 + W)
    for trial in range(400):
        alpha = [97, 98, 99][: rng.choice([1, 2, 3])]
        merges = {}
        for a in alpha:
            for b in alpha:
                if rng.random() < rng.choice([0.5, 0.9, 1.0]):
                    merges[(a, b)] = 256 + len(merges)
        n = rng.choice([0, 1, 2, seg - 1, seg, seg + 1, 3 * seg * segs_per_tile, rng.randrange(1, 40 * seg)])
        tokens = [rng.choice(alpha) for _ in range(n)]
        if rng.random() < 0.3 and n:  # long runs of one byte
            tokens = [alpha[0]] * n
        chunk = rng.choice([1 << 30, 1 << 30, 7, seg * segs_per_tile, 3 * seg * segs_per_tile + 1, 5])
        every = rng.choice([1, 2, 3, 7, 1000])
        got = run_tiled(tokens, merges, chunk, seg, segs_per_tile, W, every)
        want = []
        for s in range(0, n, chunk):
            want += pm.bpe_sweep(tokens[s:s + chunk], merges)[0]
        assert got == want, (tokens, merges, chunk, every)


def test_start_bits_trick_exhaustive():
    seg = 10
    for m in range(1 << seg):
        for cin in (0, 1):
            st, prev = 0, cin
            for j in range(seg):
                s = ((m >> j) & 1) & (1 - prev)
                st |= s << j
                prev = s
            assert start_bits(m, cin, seg) == st


# ==================================================================================================
# v2 algebra: partial ("speculated phase") tile functions and the 3-valued look-back fold.
# A tile that looked only at the pairs starting at positions of parity p (its predicted carry_in) and
# found them all mergeable knows f(p) = (carry_out p, T/2 tokens) and nothing about f(1-p).
# ==================================================================================================

def fn_identity():
    return dict(v=[1, 1], c=[0, 1], cnt=[0, 0])


def fn_compose(far, near):
    """kernels.cu compose(): carry flows far -> near; invalid (unknown) branches stay invalid."""
    out = dict(v=[0, 0], c=[0, 0], cnt=[0, 0])
    for b in (0, 1):
        m = far["c"][b]
        out["v"][b] = far["v"][b] & near["v"][m]
        out["c"][b] = near["c"][m]
        out["cnt"][b] = far["cnt"][b] + near["cnt"][m]
    return out


def desc_full(tile: "Tile"):
    ident = tile.tile_id
    return dict(state="A", v=[1, 1], c=[0 if ident else tile.tile_const, 1 if ident else tile.tile_const],
                cnt=[tile.total0, tile.total0 - tile.f_delta])


def desc_part(p, tile_elems):
    d = dict(state="A", v=[0, 0], c=[0, 0], cnt=[0, 0])
    d["v"][p], d["c"][p], d["cnt"][p] = 1, p, tile_elems // 2
    return d


def lookback_v2(status, tile, W):
    """Windowed fold toward the nearest PREFIX; returns (carry_in, base) or None if a needed branch
    of some predecessor is unknown."""
    running = fn_identity()
    j = tile - 1
    while True:
        lanes = [status[j - l] if j - l >= 0 else dict(state="P", c=[0, 0], count=0) for l in range(W)]
        p = next((l for l, s in enumerate(lanes) if s["state"] == "P"), W)
        comp = fn_identity()   # composite of lanes (W-1 .. 0), entries beyond the PREFIX masked to identity
        for l in range(W - 1, -1, -1):
            if l > p:
                e = fn_identity()
            elif l == p:
                e = dict(v=[1, 1], c=list(lanes[l]["c"]), cnt=[0, 0])
            else:
                e = lanes[l]
            comp = fn_compose(comp, e)
        total = fn_compose(comp, running)
        if p < W:
            if not total["v"][0]:
                return None
            return total["c"][0], lanes[p]["count"] + total["cnt"][0]
        running = total
        j -= W


def run_tiled_v2(tokens, merges, chunk, seg, segs_per_tile, W, rng):
    n = len(tokens)
    m_all = [0] * n
    for i in range(n - 1):
        if (i + 1) % chunk != 0 and (tokens[i], tokens[i + 1]) in merges:
            m_all[i] = 1
    values = [merges.get((tokens[i], tokens[i + 1]), None) if i + 1 < n else None for i in range(n)]
    tile_elems = seg * segs_per_tile
    n_tiles = (n + tile_elems - 1) // tile_elems
    status, truth, out = [], [], []
    hint = 0
    stats = dict(part=0, upgraded=0, full=0)
    for t in range(n_tiles):
        base_pos = t * tile_elems
        ms, vms = [], []
        for s in range(segs_per_tile):
            g = base_pos + s * seg
            ms.append(sum(m_all[g + j] << j for j in range(seg) if g + j < n))
            vms.append(sum(1 << j for j in range(seg) if g + j < n))
        tile = Tile(ms, vms, seg)
        # the real hint is stale (written by whichever tile published last): sometimes wrong
        p_hat = 0 if base_pos % chunk == 0 else (hint if rng.random() < 0.7 else rng.randrange(2))
        want = sum(1 << j for j in range(p_hat, seg, 2))
        part_ok = all((m & want) == want and vm == (1 << seg) - 1 for m, vm in zip(ms, vms))
        mode = "part" if part_ok else "full"
        desc = desc_part(p_hat, tile_elems) if part_ok else desc_full(tile)
        truth.append(desc_full(tile))
        while True:
            r = None if t else (0, 0)
            while r is None:
                r = lookback_v2(status, t, W)
                if r is None:
                    # in-flight predecessors whose speculated branch was wrong upgrade themselves
                    for k in range(t):
                        if status[k]["state"] == "A" and status[k]["v"] != [1, 1]:
                            status[k] = truth[k]
                            stats["upgraded"] += 1
            cin, base = r
            if mode == "part" and cin != p_hat:
                mode, desc = "full", desc_full(tile)
                stats["upgraded"] += 1
                continue
            break
        assert base == len(out), (t, base, len(out))
        stats[mode] += 1
        if mode == "part":   # fast emit: every segment emits the seg/2 merged ids of its parity-p pairs
            toks = [values[base_pos + j] for j in range(p_hat, tile_elems, 2)]
            assert None not in toks
            c_out = p_hat
        else:
            toks = tile.emit(cin, tokens[base_pos:] + [0] * tile_elems, values[base_pos:] + [None] * tile_elems)
            c_out = cin if tile.tile_id else tile.tile_const
        out += toks
        hint = c_out
        # mimic in-flight tiles: some predecessors are still AGGREGATEs (partial ones stay partial)
        if rng.random() < 0.5:
            status.append(dict(state="P", c=[c_out, c_out], count=base + len(toks)))
        else:
            status.append(desc if mode == "part" else desc_full(tile))
            # everything older than a few windows has certainly finished
        for k in range(max(0, t - 3 * W)):
            if status[k]["state"] != "P":
                d = truth[k]
                # recompute its prefix from the ground truth
                status[k] = dict(state="P", c=[status_c_out[k]] * 2, count=status_incl[k])
        status_c_out.append(c_out)
        status_incl.append(base + len(toks))
    return out, stats


status_c_out, status_incl = [], []


@pytest.mark.parametrize("seg,segs_per_tile,W", [(4, 2, 2), (4, 4, 3), (16, 2, 4), (8, 3, 32), (2, 2, 2)])
def test_tiled_sweep_v2_partial_functions(seg, segs_per_tile, W):
    rng = random.Random(7 * seg + segs_per_tile + W)
    seen_part = seen_up = 0
    for trial in range(400):
        alpha = [97, 98, 99][: rng.choice([1, 2, 3])]
        merges = {}
        dens = rng.choice([0.5, 0.9, 1.0, 1.0])
        for a in alpha:
            for b in alpha:
                if rng.random() < dens:
                    merges[(a, b)] = 256 + len(merges)
        n = rng.choice([0, 1, seg, seg + 1, 3 * seg * segs_per_tile, rng.randrange(1, 60 * seg)])
        tokens = [rng.choice(alpha) for _ in range(n)]
        if rng.random() < 0.4 and n:
            tokens = [alpha[0]] * n
            if rng.random() < 0.5 and n > 3:   # one break early: the rest of the run is in odd phase
                tokens[rng.randrange(0, min(n, 9))] = 120
        chunk = rng.choice([1 << 30, 1 << 30, 7, seg * segs_per_tile, 2 * seg * segs_per_tile, 3 * seg * segs_per_tile + 1, 5])
        status_c_out.clear()
        status_incl.clear()
        got, stats = run_tiled_v2(tokens, merges, chunk, seg, segs_per_tile, W, rng)
        want = []
        for s in range(0, n, chunk):
            want += pm.bpe_sweep(tokens[s:s + chunk], merges)[0]
        assert got == want, (tokens, merges, chunk)
        seen_part += stats["part"]
        seen_up += stats["upgraded"]
    assert seen_part > 50 and seen_up > 5   # both the fast path and the upgrade path were exercised
