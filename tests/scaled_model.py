"""Executable specification (numpy, any float dtype) of the arithmetic the CUDA kernels perform:
the *real-domain, power-of-two-rescaled* forward-backward that replaces the reference's
log-domain recursions ("Numerics").  It is not an oracle -- the oracle is
``oracle/asg_oracle.py`` -- it is the kernels' maths written down once so that
``tests/test_scaled_model.py`` can show, on CPU and in float32, that this formulation meets the
1e-4 parity bar against the float64 oracle before any GPU time is spent.

Per utterance (index n dropped), with R = max finite A, E = exp(A - R), m_t = max_l x[t,l],
u_t = exp(x[t] - m_t):

  FCC alpha:  P_0 = u_0;  P_t = s_t * u_t o (E P_{t-1}),   s_t = 2^-floor(log2 max(.))
  FCC beta :  Q_{T-1} = 1;  Q_t = s'_t * E^T (u_{t+1} o Q_{t+1})
  score    :  log(sum u_0 o Q_0) + sum_t m_t + (T-1) R - log(prod s'_t)
  grads    :  gamma_t = P_t o Q_t / Z_t ;  xi_t[i,j] = u_t[i] Q_t[i] E[i,j] P_{t-1}[j] s_t / Z_t
and the bidiagonal analogue for FAC with h_s = exp(A[y_s,y_s]-Rb), d_s = exp(A[y_{s+1},y_s]-Rb).
FAC states outside the feasible band (tech_report.tex:214: s <= t and S-1-s <= T-1-t) carry no
complete path; they are set to exactly 0 so the per-frame rescale follows the band's maximum.

This model renormalises every block on every frame.  The fp32 kernels do it lazily (only when a block maximum leaves
[2^-40, 4) or the inflow from the neighbour would swamp the block) and record the drift in the exponent word: the same values
up to exact powers of two.  Range rules shared with the kernels: a value more than the format's range (2^126 in fp32, 2^1022 in
fp64) below the scale it is brought to is FLUSHED to zero, never clamped to the smallest normal number; a block re-bases on what
flows in from its neighbour when that inflow would be >= 4 in the block's own units (or the block is empty).
"""
import numpy as np


def _pow2_scale(M, dt):
    """scale = 2^-floor(log2 M) (exact), exponent e; M==0 -> (1, 0)."""
    M = np.asarray(M, dtype=np.float64)
    with np.errstate(divide="ignore"):
        e = np.where(M > 0, np.floor(np.log2(np.where(M > 0, M, 1.0))), 0.0)
    return np.exp2(-e).astype(dt), e


def _emissions(x, dt):
    m = x.max(-1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0).astype(dt)
    with np.errstate(over="ignore"):
        u = np.exp((x - m).astype(dt)).astype(dt)
    return u, m[..., 0]


BLK = 8   # FAC states per block (= per GPU lane)


def asg_scaled(x, y, A, il, tl, g_full, g_aligned, dtype=np.float32):
    dt = dtype
    x = np.asarray(x).astype(dt)
    A = np.asarray(A).astype(dt)
    T, N, L = x.shape
    S = y.shape[1]
    finiteA = A[np.isfinite(A)]
    R = dt(finiteA.max()) if finiteA.size else dt(0)
    E = np.where(np.isneginf(A), 0, np.exp((A - R).astype(dt))).astype(dt)
    u, mx = _emissions(x, dt)
    ln2 = np.log(2.0)

    full = np.zeros(N); aligned = np.zeros(N)
    gx = np.zeros((T, N, L), dtype=dt)
    G = np.zeros((L, L), dtype=np.float64)        # sum of w (x) p, multiplied by E at the end
    gA_al = np.zeros((L, L), dtype=np.float64)
    for n in range(N):
        Tn, Sn = int(il[n]), int(tl[n])
        # ---------------- FCC ----------------
        P = np.zeros((Tn, L), dtype=dt); sa = np.ones(Tn, dtype=dt)
        v = u[0, n].copy()
        s, _ = _pow2_scale(v.max(), dt); P[0] = v * s; sa[0] = s
        for t in range(1, Tn):
            v = (E @ P[t - 1]).astype(dt) * u[t, n]
            s, _ = _pow2_scale(v.max(), dt)
            P[t] = v * s; sa[t] = s
        Q = np.zeros((Tn, L), dtype=dt)
        Q[Tn - 1] = 1
        esum = 0.0; msum = 0.0
        for t in range(Tn - 2, -1, -1):
            W = u[t + 1, n] * Q[t + 1]
            z = (E.T @ W).astype(dt)
            s, e = _pow2_scale(z.max(), dt)
            Q[t] = z * s
            esum += e; msum += float(mx[t + 1, n])
        with np.errstate(divide="ignore"):
            full[n] = np.log(float((u[0, n] * Q[0]).sum())) + float(mx[0, n]) + msum + (Tn - 1) * float(R) + esum * ln2
        gam = P * Q
        Z = gam.sum(-1)
        ok = Z > 0
        Zs = np.where(ok, Z, 1)
        gx[:Tn, n] += (g_full[n] * gam / Zs[:, None] * ok[:, None]).astype(dt)
        for t in range(1, Tn):
            if ok[t]:
                w = (g_full[n] * u[t, n] * Q[t] * sa[t] / Zs[t]).astype(dt)
                G += np.outer(w, P[t - 1])
        # ---------------- FAC ----------------
        ys = y[n, :Sn]
        H = A[ys, ys]
        D = A[ys[1:], ys[:-1]] if Sn > 1 else np.zeros(0, dtype=dt)
        allv = np.concatenate([H, D]); fin = allv[np.isfinite(allv)]
        Rb = dt(fin.max()) if fin.size else dt(0)
        h = np.where(np.isneginf(H), 0, np.exp((H - Rb).astype(dt))).astype(dt)
        d = np.where(np.isneginf(D), 0, np.exp((D - Rb).astype(dt))).astype(dt)
        ub = u[:Tn, n][:, ys]                         # [Tn, Sn]
        # block floating point: states are grouped in blocks of BLK (one GPU lane each); every block
        # carries its own power-of-two exponent, so true value = mantissa * 2^exponent.
        nb = (Sn + BLK - 1) // BLK
        Sp = nb * BLK
        hp = np.zeros(Sp, dtype=dt); hp[:Sn] = h
        dp = np.zeros(Sp, dtype=dt); dp[:Sn - 1] = d            # dp[s]: edge s -> s+1
        ubp = np.zeros((Tn, Sp), dtype=dt); ubp[:, :Sn] = ub
        blk_of = np.arange(Sp) // BLK

        def renorm(v, ex):
            """per-block rescale of mantissas v (Sp,) with exponents ex (nb,)"""
            vb = v.reshape(nb, BLK)
            sc, e = _pow2_scale(vb.max(1), dt)
            return (vb * sc[:, None]).reshape(Sp).astype(dt), ex + e.astype(np.int64)

        lim = 126 if dt == np.float32 else 1022          # the format's normal range in powers of two (kernel: Num<T>::kMaxShift)

        def shift(val, de):
            """val * 2^de the way the kernels do it: more than the format's range below is FLUSHED to zero (never clamped to the
            smallest normal number: that would inject mass), above it saturates at 2^lim."""
            de = np.asarray(de)
            f = np.where(de < -lim, 0.0, np.exp2(np.clip(de, -lim, lim)))
            return (val * f.astype(dt)).astype(dt)

        def rebase(prev, ex, e_nb, v_nb, b):
            """Block b receives v_nb * 2^e_nb from its neighbour.  An empty block adopts the exponent of what arrives; a block whose
            inflow would be >= 4 in its own units is re-based on the inflow (its own mantissas shrink, flushed below the range)."""
            if v_nb <= 0:
                return
            blk = prev[b * BLK:(b + 1) * BLK]
            nva = float(v_nb) * 2.0 ** float(np.clip(e_nb - ex[b], -lim, lim))
            if blk.max() == 0 or nva >= 4.0:
                en = int(np.floor(np.log2(float(v_nb))))
                En = int(e_nb) + en
                prev[b * BLK:(b + 1) * BLK] = shift(blk, np.full(BLK, min(int(ex[b]) - En, 0)))
                ex[b] = En

        Pa = np.zeros((Tn, Sp), dtype=dt); Ea = np.zeros((Tn, nb), dtype=np.int64)
        v = np.zeros(Sp, dtype=dt); v[0] = ubp[0, 0]
        Pa[0], Ea[0] = renorm(v, np.zeros(nb, dtype=np.int64))
        for t in range(1, Tn):
            ex = Ea[t - 1].copy()
            prev = Pa[t - 1].copy()
            for b in range(1, nb):                                   # adopt / re-base on what flows in from the left block
                rebase(prev, ex, Ea[t - 1][b - 1], Pa[t - 1][b * BLK - 1], b)
            left = np.zeros(Sp, dtype=dt)
            left[1:] = prev[:-1]
            left[BLK::BLK] = Pa[t - 1][BLK - 1:Sp - 1:BLK]           # block-boundary elements come from the left block AS IT WAS
            de = np.zeros(Sp, dtype=np.int64)
            de[BLK::BLK] = Ea[t - 1][:-1] - ex[1:]
            left = shift(left, de)
            v = ubp[t] * (hp * prev + np.concatenate([[0], dp[:-1]]).astype(dt) * left)
            v[:max(0, (Sn - 1) - (Tn - 1 - t))] = 0                  # band: states that can no longer reach the end
            Pa[t], Ea[t] = renorm(v.astype(dt), ex)
        Qa = np.zeros((Tn, Sp), dtype=dt); Eb = np.zeros((Tn, nb), dtype=np.int64)
        Qa[Tn - 1, Sn - 1] = 1
        msum = 0.0
        for t in range(Tn - 2, -1, -1):
            W0 = (ubp[t + 1] * Qa[t + 1]).astype(dt)
            W = W0.copy()
            ex = Eb[t + 1].copy()
            for b in range(nb - 2, -1, -1):                          # adopt / re-base on what flows in from the right block
                rebase(W, ex, Eb[t + 1][b + 1], W0[(b + 1) * BLK], b)
            right = np.zeros(Sp, dtype=dt)
            right[:-1] = W[1:]
            right[BLK - 1:Sp - 1:BLK] = W0[BLK::BLK]                 # block-boundary elements come from the right block AS IT WAS
            de = np.zeros(Sp, dtype=np.int64)
            de[BLK - 1:Sp - 1:BLK] = Eb[t + 1][1:] - ex[:-1]
            right = shift(right, de)
            z = hp * W + dp * right
            z[t + 1:] = 0                                            # band: states not reachable from (0,0) by frame t
            Qa[t], Eb[t] = renorm(z.astype(dt), ex)
            msum += float(mx[t + 1, n])
        with np.errstate(divide="ignore"):
            aligned[n] = (np.log(float(ubp[0, 0] * Qa[0, 0])) + float(mx[0, n]) + msum + (Tn - 1) * float(Rb)
                          + float(Eb[0, 0]) * ln2)
        C = Ea + Eb                                                  # [Tn, nb]
        prod = (Pa * Qa).reshape(Tn, nb, BLK)
        live = prod.max(2) > 0
        Cmax = np.where(live, C, -10**9).max(1)                      # [Tn]
        dC = C - Cmax[:, None]
        wgt = np.where(live & (dC >= -lim), np.exp2(np.clip(dC, -lim, 0)), 0).astype(dt)        # flushed, not clamped
        gam = (prod * wgt[:, :, None]).reshape(Tn, Sp)
        Z = gam.sum(-1); ok = Z > 0; Zs = np.where(ok, Z, 1)
        post = (g_aligned[n] * gam / Zs[:, None] * ok[:, None]).astype(dt)   # [Tn,Sp]
        for s_ in range(Sn):
            gx[:Tn, n, ys[s_]] += post[:, s_]
        stay = hp[None, :] * Pa[:-1]                                # into state s at t (t>=1)
        leftm = np.zeros((Tn - 1, Sp), dtype=dt); leftm[:, 1:] = Pa[:-1, :-1]
        de = np.zeros((Tn - 1, Sp), dtype=np.int64)
        de[:, BLK::BLK] = Ea[:-1, :-1] - Ea[:-1, 1:]
        leftm = shift(leftm, de)
        move = np.concatenate([[0], dp[:-1]]).astype(dt)[None, :] * leftm
        den = stay + move
        with np.errstate(invalid="ignore", divide="ignore"):
            frac = np.where(den > 0, stay / den, 0)        # share of the mass that stayed; the rest moved (as the kernels do:
            fracm = 1.0 - frac                              # the move share is what is left, never a second quotient)
        gH = (post[1:] * frac).sum(0)
        gD = (post[1:] * fracm).sum(0)                              # gD[s] = edge (s-1 -> s)
        for s_ in range(Sn):
            gA_al[ys[s_], ys[s_]] += gH[s_]
            if s_ >= 1:
                gA_al[ys[s_], ys[s_ - 1]] += gD[s_]
    gA = E.astype(np.float64) * G + gA_al
    return {"full": full, "aligned": aligned, "loss": full - aligned, "gx": gx, "gA": gA}
