"""NumPy model of the RESIDENT kernel's three-level tridiagonal solve
(thread: 8 rows, warp: 32 separators by PCR, CTA: W warp separators by Thomas).
Development aid: validates the algebra of quantpde_b200/csrc/qpde_resident.cu
against a plain Thomas solve.  Not part of the product or of the tests."""
import numpy as np

M = 8


def thomas(a, b, c, r):
    n = len(b)
    cp = np.zeros(n); dp = np.zeros(n)
    cp[0] = c[0] / b[0]; dp[0] = r[0] / b[0]
    for i in range(1, n):
        den = b[i] - a[i] * cp[i - 1]
        cp[i] = c[i] / den
        dp[i] = (r[i] - a[i] * dp[i - 1]) / den
    x = np.zeros(n)
    x[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


def shfl_up(v, d):     # value from lane - d (0 where out of the warp)
    out = np.zeros_like(v)
    W = v.shape[0] // 32
    vv = v.reshape(W, 32)
    oo = out.reshape(W, 32)
    oo[:, d:] = vv[:, :32 - d]
    return out


def shfl_down(v, d):
    out = np.zeros_like(v)
    W = v.shape[0] // 32
    vv = v.reshape(W, 32)
    oo = out.reshape(W, 32)
    oo[:, :32 - d] = vv[:, d:]
    return out


class Factor:
    pass


def factor(a, b, c):
    """a, b, c: [P, M] rows; c of the last real row must already be 0 and a of
    row 0 must be 0.  Returns the cached factorisation."""
    P = a.shape[0]
    W = P // 32
    f = Factor()
    f.a, f.b, f.c = a, b, c
    # ---- level 1: interior rows 1..7 of every thread
    m = np.zeros((P, M)); invd = np.zeros((P, M))
    d = b[:, 1].copy(); invd[:, 1] = 1. / d
    for j in range(2, M):
        m[:, j] = a[:, j] * invd[:, j - 1]
        d = b[:, j] - m[:, j] * c[:, j - 1]
        invd[:, j] = 1. / d
    Vs = np.zeros((P, M)); Ws = np.zeros((P, M))
    y = np.zeros((P, M)); y[:, 1] = a[:, 1]
    for j in range(2, M):
        y[:, j] = -m[:, j] * y[:, j - 1]
    Vs[:, M - 1] = y[:, M - 1] * invd[:, M - 1]
    Ws[:, M - 1] = c[:, M - 1] * invd[:, M - 1]
    for j in range(M - 2, 0, -1):
        Vs[:, j] = (y[:, j] - c[:, j] * Vs[:, j + 1]) * invd[:, j]
        Ws[:, j] = (0. - c[:, j] * Ws[:, j + 1]) * invd[:, j]
    f.m, f.invd, f.Vs, f.Ws = m, invd, Vs, Ws
    # ---- reduced rows (separator = row 0 of each thread)
    Vs7_prev = shfl_up(Vs[:, M - 1], 1); Ws7_prev = shfl_up(Ws[:, M - 1], 1)
    lane = np.arange(P) % 32
    A = -a[:, 0] * Vs7_prev
    B = b[:, 0] - a[:, 0] * Ws7_prev - c[:, 0] * Vs[:, 1]
    Cc = -c[:, 0] * Ws[:, 1]
    # ---- level 2: PCR over lanes 1..31 of each warp (lane 0 masked out)
    interior = lane >= 1
    A2 = np.where(interior, A, 0.); B2 = np.where(interior, B, 1.); C2 = np.where(interior, Cc, 0.)
    V2 = np.where(lane == 1, A, 0.)      # rhs A_1 e_1 (coupling to the warp separator)
    W2 = np.where(lane == 31, Cc, 0.)    # rhs C_31 e_31 (coupling to the next warp's separator)
    A2 = np.where(lane == 1, 0., A2); C2 = np.where(lane == 31, 0., C2)
    f.al, f.ga = [], []
    for s in range(5):
        dlt = 1 << s
        Bm = shfl_up(B2, dlt); Bp = shfl_down(B2, dlt)
        okm = lane - dlt >= 1; okp = lane + dlt <= 31
        al = np.where(okm & interior, -A2 / np.where(okm, Bm, 1.), 0.)
        ga = np.where(okp & interior, -C2 / np.where(okp, Bp, 1.), 0.)
        nA = al * shfl_up(A2, dlt); nC = ga * shfl_down(C2, dlt)
        nB = B2 + al * shfl_up(C2, dlt) + ga * shfl_down(A2, dlt)
        V2 = V2 + al * shfl_up(V2, dlt) + ga * shfl_down(V2, dlt)
        W2 = W2 + al * shfl_up(W2, dlt) + ga * shfl_down(W2, dlt)
        A2, B2, C2 = nA, nB, nC
        f.al.append(al); f.ga.append(ga)
    f.invB2 = 1. / B2
    f.V2 = V2 * f.invB2; f.W2 = W2 * f.invB2
    # ---- level 3: W x W tridiagonal over the warp separators
    t0 = np.arange(W) * 32
    # quantities from the previous warp's lane 31
    Vs7p = np.concatenate([[0.], Vs[t0[1:] - 1, M - 1]]); Ws7p = np.concatenate([[0.], Ws[t0[1:] - 1, M - 1]])
    V2p = np.concatenate([[0.], f.V2[t0[1:] - 1]]); W2p = np.concatenate([[0.], f.W2[t0[1:] - 1]])
    At = -a[t0, 0] * Vs7p
    Bt = b[t0, 0] - a[t0, 0] * Ws7p - c[t0, 0] * Vs[t0, 1]
    Ct = -c[t0, 0] * Ws[t0, 1]
    f.q1 = -a[t0, 0]; f.q2 = -At; f.q3 = -Ct
    f.A3 = -At * V2p
    f.B3 = Bt - At * W2p - Ct * f.V2[t0 + 1]
    f.C3 = -Ct * f.W2[t0 + 1]
    return f


def solve(f, r):
    a, b, c = f.a, f.b, f.c
    P = a.shape[0]; W = P // 32
    lane = np.arange(P) % 32
    # level 1 interior solve
    G = np.zeros((P, M))
    y = np.zeros((P, M)); y[:, 1] = r[:, 1]
    for j in range(2, M):
        y[:, j] = r[:, j] - f.m[:, j] * y[:, j - 1]
    G[:, M - 1] = y[:, M - 1] * f.invd[:, M - 1]
    for j in range(M - 2, 0, -1):
        G[:, j] = (y[:, j] - c[:, j] * G[:, j + 1]) * f.invd[:, j]
    G7p = shfl_up(G[:, M - 1], 1)
    R = r[:, 0] - a[:, 0] * G7p - c[:, 0] * G[:, 1]
    R2 = np.where(lane >= 1, R, 0.)
    for s in range(5):
        dlt = 1 << s
        R2 = R2 + f.al[s] * shfl_up(R2, dlt) + f.ga[s] * shfl_down(R2, dlt)
    G2 = R2 * f.invB2
    # level 3
    t0 = np.arange(W) * 32
    G7prev = np.concatenate([[0.], G[t0[1:] - 1, M - 1]])
    G2prev = np.concatenate([[0.], G2[t0[1:] - 1]])
    P0 = r[t0, 0] - c[t0, 0] * G[t0, 1]
    R3 = P0 + f.q1 * G7prev + f.q2 * G2prev + f.q3 * G2[t0 + 1]
    Y = thomas(f.A3, f.B3, f.C3, R3)
    Yn = np.concatenate([Y[1:], [0.]])
    # back substitution
    w = np.arange(P) // 32
    ysep = np.where(lane == 0, Y[w], G2 - Y[w] * f.V2 - Yn[w] * f.W2)
    ynext = shfl_down(ysep, 1)
    ynext = np.where(lane == 31, Yn[w], ynext)
    x = np.zeros((P, M))
    x[:, 0] = ysep
    for j in range(1, M):
        x[:, j] = G[:, j] - ysep * f.Vs[:, j] - ynext * f.Ws[:, j]
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, P in ((2048, 256), (2112, 288), (500, 64), (33, 32)):
        # a CN-like M-matrix with weak diagonal dominance and a penalised prefix
        al = rng.uniform(50, 500, n); be = rng.uniform(50, 500, n)
        a = -al; c = -be; b = 1. + al + be + 0.02
        a[0] = 0.; c[-1] = 0.
        b[: n // 3] += 1e6
        r = rng.uniform(0, 100, n)
        ap = np.zeros(P * M); bp = np.ones(P * M); cp = np.zeros(P * M); rp = np.zeros(P * M)
        ap[:n] = a; bp[:n] = b; cp[:n] = c; rp[:n] = r
        f = factor(ap.reshape(P, M), bp.reshape(P, M), cp.reshape(P, M))
        x = solve(f, rp.reshape(P, M)).reshape(-1)[:n]
        xr = thomas(a, b, c, r)
        print(n, P, np.abs(x - xr).max() / np.abs(xr).max())
