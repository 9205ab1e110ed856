"""Lane-level numpy model of scl_decode_kernel (polar_b200/csrc/polar_b200.cu).

TEST INFRASTRUCTURE: it re-states the kernel's *organisation* -- butterfly-order layer
arrays stored [beta][lane], 5-bit column pointers instead of lazy copies, bit-packed
partial sums built in place, free-path stack held across lanes, u-hat recovered by a
packed polar transform -- in float64 numpy, one array element per warp lane, so the
design can be checked against the oracle on a machine without a GPU. It is not a
decoder anybody should call.
"""
import numpy as np


def _softplus_ref(x):
    with np.errstate(over="ignore"):
        return np.log(1.0 + np.exp(x))


def _f_rule(a, b):
    ma, mb = np.abs(a), np.abs(b)
    with np.errstate(over="ignore", invalid="ignore"):
        exact = np.log((np.exp(a + b) + 1.0) / (np.exp(a) + np.exp(b)))
    ms = np.sign(a) * np.sign(b) * np.minimum(ma, mb)
    return np.where(np.maximum(ma, mb) < 40.0, exact, ms)


def _fns(mask, k):
    """position of the k-th (1-based) set bit of mask"""
    for b in range(32):
        if (mask >> b) & 1:
            k -= 1
            if k == 0:
                return b
    return -1


def decode_group(llr, n, K, crc, frozen, order, crc_matrix, L):
    """llr: [G][N] for G = 32 // W codewords handled by one warp. Returns [G][K] bits."""
    N = 1 << n
    W = 1
    while W < L:
        W <<= 1
    G = 32 // W
    llr = np.asarray(llr, np.float64).reshape(-1, N)
    nvalid = llr.shape[0]
    assert nvalid <= G
    lane = np.arange(32)
    slot = lane & (W - 1)
    gbase = lane & ~(W - 1)
    cwi = np.minimum(lane // W, nvalid - 1)
    valid = (lane // W) < nvalid
    gmask = (1 << W) - 1
    NW = (N + 31) // 32

    X = {lam: np.zeros((1 << (n - lam), 32)) for lam in range(1, n)}
    S = {lam: np.zeros((max(1, (1 << (n - lam)) // 32), 32), np.uint64) for lam in range(0, n)}
    px = np.zeros((n + 1, 32), np.int64)
    ps = np.zeros((n + 1, 32), np.int64)
    active = valid & (slot == L - 1)
    pm = np.zeros(32)
    s_n = np.zeros(32, np.uint64)
    stk = slot.copy()
    sp = np.full(32, L - 1)
    lam_n = np.zeros(32)
    M32 = np.uint64(0xFFFFFFFF)

    def brev(i, bits):
        r = 0
        for b in range(bits):
            if i & (1 << b):
                r |= 1 << (bits - 1 - b)
        return r

    for phi in range(N):
        lam_top = 1 if phi == 0 else n - ((phi & -phi).bit_length() - 1)
        for lam in range(lam_top, n + 1):
            M = 1 << (n - lam)
            is_g = (lam == lam_top) and phi != 0
            for i in range(M):
                beta = i
                if lam == 1:
                    x0 = llr[cwi, 2 * i]
                    x1 = llr[cwi, 2 * i + 1]
                    beta = brev(i, n - 1) if n > 1 else 0
                else:
                    col = px[lam - 1]
                    x0 = X[lam - 1][i, col]
                    x1 = X[lam - 1][i + M, col]
                if is_g:
                    if lam == n:
                        bit = s_n & np.uint64(1)
                    else:
                        bit = (S[lam][beta >> 5, ps[lam]] >> np.uint64(beta & 31)) & np.uint64(1)
                    y = x1 + np.where(bit == 1, -x0, x0)
                else:
                    y = _f_rule(x0, x1)
                if lam == n:
                    lam_n = np.where(active, y, lam_n)
                else:
                    X[lam][beta, lane[active]] = y[active]
            if lam < n:
                px[lam] = lane

        u = np.zeros(32, np.uint64)
        if frozen[phi]:
            pm = np.where(active, pm + _softplus_ref(-lam_n), pm)
        else:
            m0 = pm + _softplus_ref(-lam_n)
            m1 = pm + _softplus_ref(lam_n)
            keep0 = active.copy()
            keep1 = active.copy()
            for g0 in range(0, 32, W):
                lanes = np.arange(g0, g0 + W)
                A = int(active[lanes].sum())
                if 2 * A > L:
                    forks = []
                    for l in lanes:
                        if active[l]:
                            forks.append((m0[l], 2 * (l - g0)))
                            forks.append((m1[l], 2 * (l - g0) + 1))
                    forks.sort()
                    kept = set(idx for _, idx in forks[:L])
                    for l in lanes:
                        keep0[l] = active[l] and (2 * (l - g0)) in kept
                        keep1[l] = active[l] and (2 * (l - g0) + 1) in kept
            kill = active & ~keep0 & ~keep1
            clone = keep0 & keep1
            src = lane.copy()
            tgt = np.zeros(32, np.int64)
            for g0 in range(0, 32, W):
                Kg = sum(1 << (l - g0) for l in range(g0, g0 + W) if kill[l])
                Cg = sum(1 << (l - g0) for l in range(g0, g0 + W) if clone[l])
                nk, nc = bin(Kg).count("1"), bin(Cg).count("1")
                spg = int(sp[g0])
                for l in range(g0, g0 + W):
                    s = l - g0
                    if spg <= s < spg + nk:
                        stk[l] = _fns(Kg, s - spg + 1)
                sp2 = spg + nk
                for l in range(g0, g0 + W):
                    if clone[l]:
                        ci = bin(Cg & ((1 << (l - g0)) - 1)).count("1")
                        t = stk[g0 + ((sp2 - 1 - ci) & (W - 1))]
                        src[g0 + t] = l
                sp[g0:g0 + W] = sp2 - nc
            is_new = src != lane
            new_active = active.copy()
            new_pm = pm.copy()
            for l in range(32):
                if is_new[l]:
                    s = src[l]
                    new_active[l] = True
                    new_pm[l] = m1[s]
                    u[l] = 1
                    px[:, l] = px[:, s]
                    ps[:, l] = ps[:, s]
                    s_n[l] = s_n[s]
                elif kill[l]:
                    new_active[l] = False
                    new_pm[l] = 0.0
                elif active[l]:
                    u[l] = 0 if keep0[l] else 1
                    new_pm[l] = m0[l] if keep0[l] else m1[l]
            active, pm = new_active, new_pm

        if phi % 2 == 0:
            s_n = u.copy()
        else:
            t = ((~phi) & (phi + 1)).bit_length() - 1
            lam_end = n - t
            P = u.copy()
            lam = n
            while lam > lam_end and (n - lam) < 5:
                M = 1 << (n - lam)
                Sw = s_n if lam == n else S[lam][0, ps[lam]]
                P = (((Sw ^ P) & np.uint64((1 << M) - 1)) | (P << np.uint64(M))) & M32
                lam -= 1
            if lam == lam_end:
                S[lam][0, lane] = P
            else:
                Wd = 1 << (t - 5)
                D = S[lam_end]
                D[Wd - 1, lane] = P
                while lam > lam_end:
                    mw = 1 << (n - lam - 5)
                    base = Wd - mw
                    for w in range(mw):
                        D[base - mw + w, lane] = S[lam][w, ps[lam]] ^ D[base + w, lane]
                    lam -= 1
            if lam_end >= 1:
                ps[lam_end] = lane

    D = S[0].copy()
    sw = NW >> 1
    while sw >= 1:
        for i in range(NW):
            if (i & sw) == 0:
                D[i] ^= D[i + sw]
        sw >>= 1
    for sh, msk in ((16, 0x0000FFFF), (8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
        if N > sh:
            D ^= (D >> np.uint64(sh)) & np.uint64(msk)
    uhat = np.zeros((32, N), np.uint8)
    for p in range(N):
        uhat[:, p] = ((D[p >> 5] >> np.uint64(p & 31)) & np.uint64(1)).astype(np.uint8)
    out = np.zeros((nvalid, K), np.uint8)
    for g in range(nvalid):
        g0 = g * W
        passing = []
        for l in range(g0, g0 + W):
            if not active[l]:
                continue
            ok = True
            for r in range(crc):
                acc = int((crc_matrix[r] & uhat[l, order[:K]]).sum()) & 1
                if acc != uhat[l, order[K + r]]:
                    ok = False
                    break
            if ok:
                passing.append(l)
        use_parity = crc != 0 and len(passing) > 0
        best, best_pm = g0, np.finfo(np.float64).max
        for l in range(g0, g0 + W):
            if not active[l] or (use_parity and l not in passing):
                continue
            if pm[l] < best_pm:
                best_pm, best = pm[l], l
        if active[best]:
            out[g] = uhat[best, order[:K]]
    return out
