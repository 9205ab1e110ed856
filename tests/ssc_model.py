"""numpy model of sc_ssc.cuh (TEST INFRASTRUCTURE): interprets the schedule the library builds for a code exactly the way
the kernel does -- four lanes per codeword, lane q owning quarter q of every node, rate-0 / rate-1 shortcuts, four-entry
nodes finished leaf by leaf, n butterfly stages over the re-encoded codeword and the output map -- in float64 with the
reference's literal check-node formula (PolarC/PolarCode.cpp:438-446), vectorised over codewords. What it does not model
is the kernel's word packing and memory placement."""
import ctypes as C

import numpy as np

OP_END, OP_F, OP_G, OP_G0, OP_C, OP_R0, OP_R1, OP_SUB = range(8)


def schedule(lib, n, frozen):
    frozen = np.ascontiguousarray(frozen, np.uint8)
    cnt = lib.polar_b200_ssc_schedule(n, frozen.ctypes.data, None, 0)
    if cnt <= 0:
        return None
    ops = np.zeros(cnt, np.uint32)
    assert lib.polar_b200_ssc_schedule(n, frozen.ctypes.data, ops.ctypes.data, cnt) == cnt
    return ops


def positions(lib, n, order, K):
    order = np.ascontiguousarray(order, np.uint16)
    pos = np.zeros(K, np.uint16)
    assert lib.polar_b200_ssc_positions(n, order.ctypes.data, K, pos.ctypes.data) == 0
    return pos


def f_ref(a, b):
    with np.errstate(all="ignore"):      # (the branch that overflows is the one np.where discards)
        exact = np.log((np.exp(a + b) + 1.0) / (np.exp(a) + np.exp(b)))
    sm = np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))
    return np.where(np.maximum(np.abs(a), np.abs(b)) < 40.0, exact, sm)


def g_ref(a, b, u):
    return (1.0 - 2.0 * u) * a + b


def _leaf4(al, fm, margin):
    """al: [B, 4] (lane q holds entry q); returns ([B, 4] partial sums, margin)"""
    B = al.shape[0]
    out = np.zeros((B, 4), np.uint8)
    xl = np.zeros((B, 2), np.uint8)
    xr = np.zeros((B, 2), np.uint8)
    a0, a1 = al[:, 0::2], al[:, 1::2]                        # pairs b = 0, 1
    if (fm & 3) != 3:
        l = f_ref(a0, a1)
        u0 = np.zeros(B, np.uint8); u1 = np.zeros(B, np.uint8)
        if not fm & 1:
            tt = f_ref(l[:, 0], l[:, 1]); margin = np.minimum(margin, np.abs(tt)); u0 = (tt < 0).astype(np.uint8)
        if not fm & 2:
            tt = g_ref(l[:, 0], l[:, 1], u0.astype(np.float64)); margin = np.minimum(margin, np.abs(tt)); u1 = (tt < 0).astype(np.uint8)
        xl = np.stack([u0 ^ u1, u1], axis=1)
    if (fm & 12) != 12:
        r = g_ref(a0, a1, xl.astype(np.float64))
        u2 = np.zeros(B, np.uint8); u3 = np.zeros(B, np.uint8)
        if not fm & 4:
            tt = f_ref(r[:, 0], r[:, 1]); margin = np.minimum(margin, np.abs(tt)); u2 = (tt < 0).astype(np.uint8)
        if not fm & 8:
            tt = g_ref(r[:, 0], r[:, 1], u2.astype(np.float64)); margin = np.minimum(margin, np.abs(tt)); u3 = (tt < 0).astype(np.uint8)
        xr = np.stack([u2 ^ u3, u3], axis=1)
    out[:, 0::2] = xl ^ xr
    out[:, 1::2] = xr
    return out, margin


def _sub_node(a, fm, margin):
    """a: [B, 4, E] (E entries per lane, node of 4 E leaves), fm: frozen pattern of its leaves; the kernel's sub_node<E>"""
    B, _, E = a.shape
    M = 4 * E
    full = (1 << M) - 1
    if fm == full:
        return np.zeros((B, 4, E), np.uint8), margin
    if fm == 0:
        return (a < 0).astype(np.uint8), np.minimum(margin, np.abs(a).min(axis=(1, 2)))
    if E == 1:
        out, margin = _leaf4(a[:, :, 0], fm, margin)
        return out[:, :, None], margin
    half = (1 << (M // 2)) - 1
    fl, fr = fm & half, fm >> (M // 2)
    xl = np.zeros((B, 4, E // 2), np.uint8)
    xr = np.zeros((B, 4, E // 2), np.uint8)
    a0, a1 = a[:, :, 0::2], a[:, :, 1::2]
    if fl != half:
        xl, margin = _sub_node(f_ref(a0, a1), fl, margin)
    if fr != half:
        xr, margin = _sub_node(g_ref(a0, a1, xl.astype(np.float64)), fr, margin)
    out = np.empty((B, 4, E), np.uint8)
    out[:, :, 0::2] = xl ^ xr
    out[:, :, 1::2] = xr
    return out, margin


def decode(ops, pos, n, K, llr):
    """llr: [B, N] float; returns (decoded [B, K] uint8, margin [B])."""
    llr = np.asarray(llr, np.float64)
    B, N = llr.shape
    assert N == 1 << n
    # X[lam]: [B, 4, cnt]; lane q of layer 0 = quarter q of the channel row
    X = {0: llr.reshape(B, 4, N // 4)}
    S = {}                                    # (lam, side) -> [B, 4, bits] partial sums of the node in that slot
    margin = np.full(B, np.inf)
    root = None
    pc = 0
    while True:
        op = int(ops[pc]); pc += 1
        t, m, side = op & 7, (op >> 3) & 15, (op >> 7) & 1
        lam = n - m
        if t == OP_END:
            break
        if t in (OP_F, OP_G, OP_G0, OP_R1) and lam == 1:
            # layer 1 is never stored: the root's children compute it from the channel row on the fly
            half = (op >> 8) & 3
            c0, c1 = X[0][:, :, 0::2], X[0][:, :, 1::2]
            X[1] = {1: lambda: f_ref(c0, c1), 2: lambda: g_ref(c0, c1, S[(1, 0)].astype(np.float64)), 3: lambda: c0 + c1}[half]()
        if t in (OP_F, OP_G, OP_G0):
            assert 6 <= m < n
            src = X[lam]
            a, b = src[:, :, 0::2], src[:, :, 1::2]
            if t == OP_F:
                X[lam + 1] = f_ref(a, b)
            elif t == OP_G0:
                X[lam + 1] = a + b
            else:
                X[lam + 1] = g_ref(a, b, S[(lam + 1, 0)].astype(np.float64))
        elif t == OP_C:
            cl = 1 << (m - 3)
            r = np.zeros((B, 4, cl), np.uint8) if op & (1 << 13) else S[(lam + 1, 1)]
            l = (np.zeros((B, 4, cl), np.uint8) if op & (1 << 12) else S[(lam + 1, 0)]) ^ r
            out = np.empty((B, 4, 2 * cl), np.uint8)
            out[:, :, 0::2] = l
            out[:, :, 1::2] = r
            if lam == 0:
                root = out
            else:
                S[(lam, side)] = out
        elif t == OP_R0:
            S[(lam, side)] = np.zeros((B, 4, 1 << (m - 2)), np.uint8)
        elif t == OP_R1:
            assert m >= 5
            v = X[lam]
            margin = np.minimum(margin, np.abs(v).min(axis=(1, 2)))
            S[(lam, side)] = (v < 0).astype(np.uint8)
        else:
            assert m == 5
            fm = int(ops[pc]); pc += 1
            S[(lam, side)], margin = _sub_node(X[lam], fm, margin)
    assert root is not None
    y = root.reshape(B, N).copy()                # position q N/4 + local
    s = 1
    while s < N:
        idx = np.arange(N)
        lo = idx[(idx & s) == 0]
        y[:, lo] ^= y[:, lo + s]
        s <<= 1
    e = pos.astype(np.int64)
    where = ((e >> 5) & 3) * (N // 4) + (e >> 7) * 32 + (e & 31)
    return y[:, where], margin
