"""CPU model of the carry/count algebra the CUDA sweep kernels use (blt_b200/csrc/sweep3.cuh), checked
against the reference's sequential sweep (oracle/py_model.py).

The kernels cannot run without a GPU, so this file re-derives, in plain Python with the kernels' own
variable names, every non-obvious step: the start-bit trick, identity/constant segment functions, the
"first non-identity segment" delta, the per-range carry function count_kernel accumulates, the
function composition scan_kernel performs, and emit_kernel's concrete-carry walk - including the
"dense round" shortcut (only the pairs of the carry's parity are looked at).  Small segment/range
sizes make identity segments, identity ranges and odd-phase dense runs common in random tests.
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


def range_function(ms, vms, seg):
    """count_kernel: one warp's range reduced to (identity, const, cnt0, delta)."""
    t_id, t_const, cnt0, delta = True, 0, 0, 0
    for m, vm in zip(ms, vms):
        ident, const = seg_fn(m, seg)
        cin0 = 0 if t_id else t_const
        c, _, _ = seg_count(m, vm, cin0, seg)
        cnt0 += c
        if not ident:
            if t_id:  # the first non-identity segment is the only one whose count sees the range's carry_in
                c1, _, _ = seg_count(m, vm, 1, seg)
                delta = c - c1
                assert delta in (0, 1)
            t_id, t_const = False, const
    return dict(id=t_id, cst=t_const, delta=delta, cnt0=cnt0)


def scan_compose(far, near):
    """sweep3.cuh scan_compose(): carry flows far -> near."""
    c_mid0 = 0 if far["id"] else far["cst"]
    return dict(cnt0=far["cnt0"] + near["cnt0"] - (1 if (c_mid0 and near["delta"]) else 0),
                delta=near["delta"] if far["id"] else far["delta"],
                cst=far["cst"] if near["id"] else near["cst"],
                id=far["id"] and near["id"])


def emit_range(ms, vms, seg, carry, tokens, values, base_pos, dense_stats):
    """emit_kernel: walk the range with a concrete carry.  A segment whose pairs of the carry's parity are
    all rules is 'dense': only those pairs are looked at and the carry is unchanged."""
    out = []
    for i, (m, vm) in enumerate(zip(ms, vms)):
        full = vm == (1 << seg) - 1
        par = sum(1 << j for j in range(carry, seg, 2))
        if full and (m & par) == par:
            k0 = base_pos + i * seg
            out += [values[k0 + j] for j in range(carry, seg, 2)]
            dense_stats[0] += 1
            continue
        cnt, st, em = seg_count(m, vm, carry, seg)
        for j in range(seg):
            if (em >> j) & 1:
                k = base_pos + i * seg + j
                out.append(values[k] if (st >> j) & 1 else tokens[k])
        ident, const = seg_fn(m, seg)
        if not ident:
            carry = const
    return out, carry


def run_three_kernels(tokens, merges, chunk, seg, segs_per_range):
    n = len(tokens)
    m_all = [0] * n
    for i in range(n - 1):
        if (i + 1) % chunk != 0 and (tokens[i], tokens[i + 1]) in merges:
            m_all[i] = 1
    values = [merges.get((tokens[i], tokens[i + 1]), None) if (i + 1 < n and m_all[i]) else None for i in range(n)]
    r_elems = seg * segs_per_range
    n_ranges = (n + r_elems - 1) // r_elems
    ranges = []
    for t in range(n_ranges):
        ms, vms = [], []
        for s in range(segs_per_range):
            g = t * r_elems + s * seg
            ms.append(sum(m_all[g + j] << j for j in range(seg) if g + j < n))
            vms.append(sum(1 << j for j in range(seg) if g + j < n))
        ranges.append((ms, vms))
    fns = [range_function(ms, vms, seg) for ms, vms in ranges]                  # count_kernel
    out, carry, base, dense = [], 0, 0, [0]
    prefix = dict(id=True, cst=0, delta=0, cnt0=0)
    for t, (ms, vms) in enumerate(ranges):                                      # scan_kernel + emit_kernel
        c_in = 0 if prefix["id"] else prefix["cst"]                             # the launch starts with carry 0
        assert c_in == carry and prefix["cnt0"] == base == len(out), (t, c_in, carry, prefix, base, len(out))
        toks, carry = emit_range(ms, vms, seg, carry, tokens + [0] * r_elems, values + [None] * r_elems,
                                 t * r_elems, dense)
        assert None not in toks
        out += toks
        base += fns[t]["cnt0"] - (fns[t]["delta"] if c_in else 0)
        prefix = scan_compose(prefix, fns[t])
    return out, dense[0]


@pytest.mark.parametrize("seg,segs_per_range", [(4, 2), (4, 5), (16, 2), (8, 3), (2, 2)])
def test_count_scan_emit_equals_sequential(seg, segs_per_range):
    rng = random.Random(seg * 1000 + segs_per_range)
    dense_hits = 0
    for trial in range(500):
        alpha = [97, 98, 99][: rng.choice([1, 2, 3])]
        merges = {}
        dens = rng.choice([0.5, 0.9, 1.0, 1.0])
        for a in alpha:
            for b in alpha:
                if rng.random() < dens:
                    merges[(a, b)] = 256 + len(merges)
        n = rng.choice([0, 1, 2, seg - 1, seg, seg + 1, 3 * seg * segs_per_range, rng.randrange(1, 60 * seg)])
        tokens = [rng.choice(alpha) for _ in range(n)]
        if rng.random() < 0.4 and n:  # long runs of one byte ...
            tokens = [alpha[0]] * n
            if rng.random() < 0.5 and n > 3:   # ... with one early break, so the rest of the run is in odd phase
                tokens[rng.randrange(0, min(n, 9))] = 120
        chunk = rng.choice([1 << 30, 1 << 30, 7, seg * segs_per_range, 2 * seg * segs_per_range,
                            3 * seg * segs_per_range + 1, 5])
        got, dense = run_three_kernels(tokens, merges, chunk, seg, segs_per_range)
        want = []
        for s in range(0, n, chunk):
            want += pm.bpe_sweep(tokens[s:s + chunk], merges)[0]
        assert got == want, (tokens, merges, chunk)
        dense_hits += dense
    assert dense_hits > 100  # the dense-round shortcut was exercised (both phases, see the odd-phase inputs)


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


def test_dense_speculation_condition():
    """dense_pairs_kernel: if every pair starting at an even offset of its chunk is a rule, the sweep's output
    is exactly those pairs' ids (plus a trailing raw byte for an odd-length chunk) - whatever the odd pairs are.
    Holds for even chunk sizes (no even pair straddles a wall)."""
    rng = random.Random(99)
    for trial in range(300):
        alpha = [97, 98, 99]
        chunk = rng.choice([2, 4, 6, 16, 64, 1 << 20])
        n = rng.randrange(0, 200)
        tokens = [rng.choice(alpha) for _ in range(n)]
        merges = {}
        for s in range(0, n, chunk):
            c = tokens[s:s + chunk]
            for i in range(0, len(c) - 1, 2):
                merges[(c[i], c[i + 1])] = 256 + len(merges) % 50
        for _ in range(rng.randrange(0, 4)):  # odd pairs may or may not be rules
            merges.setdefault((rng.choice(alpha), rng.choice(alpha)), 400)
        want = []
        for s in range(0, n, chunk):
            want += pm.bpe_sweep(tokens[s:s + chunk], merges)[0]
        got = []
        for s in range(0, n, chunk):
            c = tokens[s:s + chunk]
            got += [merges[(c[i], c[i + 1])] for i in range(0, len(c) - 1, 2)]
            if len(c) % 2:
                got.append(c[-1])
        assert got == want


def range_meta(ms, vms, seg):
    """count_kernel with a.meta: per segment (cin0, dep, d1, cnt, ident) - sweep3.cuh, 'what the walk variant
    of emit needs to know about this segment before it looks anything up'."""
    t_id, t_const, meta = True, 0, []
    for m, vm in zip(ms, vms):
        ident, const = seg_fn(m, seg)
        cin0 = 0 if t_id else t_const
        c, _, _ = seg_count(m, vm, cin0, seg)
        dep, d1 = 0, 0
        if t_id:
            dep = 1
            c1, _, _ = seg_count(m, vm, 1, seg)
            d1 = (c - c1) & 1
        meta.append(dict(cin0=cin0, dep=dep, d1=d1, cnt=c, ident=ident))
        if not ident:
            t_id, t_const = False, const
    return meta


def emit_range_walk(meta, seg, range_carry, tokens, table, base_pos):
    """emit_kernel<..., WALK>: carry_in and token count of every segment come from the meta byte; the lane then
    walks its positions and reads the table only where the scan stands (walk_step)."""
    out, carry = [], range_carry
    for i, mb in enumerate(meta):
        cin = range_carry if mb["dep"] else mb["cin0"]
        cnt = mb["cnt"] - (cin & mb["d1"] & mb["dep"])
        assert cin == carry, "the meta byte's carry_in is the carry the previous segment left"
        k0 = base_pos + i * seg
        if mb["ident"]:  # dense: every pair is a rule, merges all the way at the carry's parity
            seg_out = [table[(tokens[k0 + j], tokens[k0 + j + 1])] for j in range(cin, seg, 2)]
            standing_after = 1 - cin if seg % 2 == 0 else cin
        else:
            seg_out, standing = [], 1 - cin
            for j in range(seg):
                if standing:
                    pair = (tokens[k0 + j], tokens[k0 + j + 1])
                    hit = pair in table
                    seg_out.append(table[pair] if hit else tokens[k0 + j])
                    standing = 0 if hit else 1
                else:
                    standing = 1
            standing_after = standing
        assert len(seg_out) == cnt, (mb, cin, seg_out)
        out += seg_out
        carry = 1 - standing_after
    return out, carry


@pytest.mark.parametrize("seg,segs_per_range", [(4, 3), (16, 2), (8, 4)])
def test_walk_variant_meta_and_walk(seg, segs_per_range):
    """BLT_SWEEP_VARIANT=2 on wall-free, whole tiles (the only ones it handles itself): meta byte + walk equal
    the sequential sweep; the identity bit - not the token count - is what marks a dense segment."""
    rng = random.Random(seg * 77 + segs_per_range)
    saw_cnt_half_not_ident = 0
    for trial in range(400):
        alpha = [97, 98, 99][: rng.choice([1, 2, 3])]
        merges = {}
        dens = rng.choice([0.4, 0.8, 1.0, 1.0])
        for a in alpha:
            for b in alpha:
                if rng.random() < dens:
                    merges[(a, b)] = 256 + len(merges)
        r_elems = seg * segs_per_range
        n_ranges = rng.randrange(1, 6)
        n = n_ranges * r_elems
        tokens = [rng.choice(alpha) for _ in range(n)]
        if rng.random() < 0.4:
            tokens = [alpha[0]] * n
            if rng.random() < 0.6:
                tokens[rng.randrange(0, min(n, 9))] = 120
        want = pm.bpe_sweep(tokens, merges)[0]
        padded = tokens + [0] * (seg + 1)          # the look-ahead element of the last segment (never a rule)
        m_all = [1 if (i + 1 < n and (tokens[i], tokens[i + 1]) in merges) else 0 for i in range(n)]
        out, carry = [], 0
        prefix = dict(id=True, cst=0, delta=0, cnt0=0)
        for t in range(n_ranges):
            ms = [sum(m_all[t * r_elems + s * seg + j] << j for j in range(seg)) for s in range(segs_per_range)]
            vms = [(1 << seg) - 1] * segs_per_range
            fn, meta = range_function(ms, vms, seg), range_meta(ms, vms, seg)
            for m, mb in zip(ms, meta):
                if mb["cnt"] == seg // 2 and not mb["ident"]:
                    saw_cnt_half_not_ident += 1
            c_in = 0 if prefix["id"] else prefix["cst"]
            assert c_in == carry
            toks, carry = emit_range_walk(meta, seg, carry, padded, merges, t * r_elems)
            out += toks
            prefix = scan_compose(prefix, fn)
        # the last element of the input has no partner: the kernel never takes that tile through the walk (it is
        # not 'simple'); here the padded look-ahead 0 is simply not a rule
        assert out == want, (tokens, merges)
    assert saw_cnt_half_not_ident > 0  # seg/2 tokens without being all-rules happens: why the identity bit exists
