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
