"""numpy prototype of the exact algorithm the CUDA engine runs (packed sqrt-weight rows, Taylor-16
Paterson-Stockmeyer expm + derivative chain, gather form of G).  Used while developing; validated against
oracle/csa_oracle.py.  Not imported by the product."""
import math, sys
import numpy as np
sys.path.insert(0, "oracle")
import csa_oracle as o

C = [1.0 / math.factorial(k) for k in range(17)]

def expm_chain(K, theta_m=1.0):
    n = K.shape[0]
    norm1 = np.abs(K).sum(axis=0).max()
    s = max(0, int(math.ceil(math.log2(norm1 / theta_m)))) if norm1 > 0 else 0
    X = K * 2.0 ** (-s)
    I = np.eye(n)
    X2 = X @ X; X3 = X2 @ X; X4 = X2 @ X2
    P = [C[4*j]*I + C[4*j+1]*X + C[4*j+2]*X2 + C[4*j+3]*X3 for j in range(4)]
    Q3 = P[3] + C[16]*X4
    Q2 = P[2] + X4 @ Q3
    Q1 = P[1] + X4 @ Q2
    R = [P[0] + X4 @ Q1]
    for i in range(s):
        R.append(R[-1] @ R[-1])
    return dict(s=s, X=X, X2=X2, X3=X3, X4=X4, Q3=Q3, Q2=Q2, Q1=Q1, R=R)

def dexpm_chain(ch, E):
    s = ch["s"]; X, X2, X3, X4, Q3, Q2, Q1, R = (ch[k] for k in ("X","X2","X3","X4","Q3","Q2","Q1","R"))
    dX = E * 2.0 ** (-s)
    dX2 = dX @ X + X @ dX
    dX3 = dX2 @ X + X2 @ dX
    dX4 = dX2 @ X2 + X2 @ dX2
    dP = [C[4*j+1]*dX + C[4*j+2]*dX2 + C[4*j+3]*dX3 for j in range(4)]
    dQ3 = dP[3] + C[16]*dX4
    dQ2 = dP[2] + dX4 @ Q3 + X4 @ dQ3
    dQ1 = dP[1] + dX4 @ Q2 + X4 @ dQ2
    dR = dP[0] + dX4 @ Q1 + X4 @ dQ1
    for i in range(s):
        dR = dR @ R[i] + R[i] @ dR
    return dR

def device_algo(x, tbt_f, n, packed):
    """tbt_f: Fortran-ordered view semantic: mem[p + n q + n^2 r + n^3 s]."""
    cL = n*(n+1)//2
    lam, th = x[:cL], x[cL:]
    K = o.construct_anti_symmetric(n, th)
    ch = expm_chain(K); U = ch["R"][-1]
    L = o.get_cartan_matrix(n, lam)
    mem = np.asarray(tbt_f).reshape(-1, order="F")
    Tm = mem.reshape(n*n, n*n)            # Tm[i=(rs)][j=(pq)] = mem[j + n^2 i]   (row-major kernel view)
    if packed:
        pa = []; pq = []
        for q in range(n):
            for a in range(q+1):
                pa.append(a); pq.append(q)
        pa = np.array(pa); pq = np.array(pq)
        sw = np.where(pa == pq, 1.0, math.sqrt(2.0))
        rows = pa + n*pq
        Tt = Tm[np.ix_(rows, rows)] * sw[:, None] * sw[None, :]
    else:
        idx = np.arange(n*n); pa = idx % n; pq = idx // n; sw = np.ones(n*n)
        Tt = Tm
    At = (U[pa, :] * U[pq, :]) * sw[:, None]
    Bt = At @ L
    D = Bt @ At.T - Tt
    cost = np.sum(D*D)
    Et = D @ At
    S = At.T @ Et
    F = Et @ L
    G = np.zeros((n, n))
    if packed:
        pk = lambda a, q: (max(a, q)*(max(a, q)+1)//2 + min(a, q))
        for a in range(n):
            for q in range(n):
                r = pk(a, q)
                G[a, :] += 4.0 / sw[r] * U[q, :] * F[r, :]
    else:
        for a in range(n):
            for q in range(n):
                G[a, :] += 2.0 * U[q, :] * (F[a + n*q, :] + F[q + n*a, :])
    Lt = dexpm_chain(ch, G.T)
    gl = np.array([2.0*S[i, j]*(2.0 if i != j else 1.0) for i in range(n) for j in range(i+1)])
    gt = np.array([2.0*(Lt[q, p] - Lt[p, q]) for p in range(n) for q in range(p+1, n)])
    return cost, np.concatenate([gl, gt])

if __name__ == "__main__":
    for n in (4, 7, 12):
        x = o.random_x(n, 3)
        for kind, packed in ((o.dense_sym_tbt, True), (o.dense_sym_tbt, False), (o.pairsym_tbt, False)):
            T = kind(n, 1)
            c0, g0 = o.cost_grad_factored(x, T)
            c1, g1 = device_algo(x, np.asfortranarray(T), n, packed)
            print(n, kind.__name__, packed, abs(c1-c0)/abs(c0), np.max(np.abs(g1-g0))/np.max(np.abs(g0)))
    # large-norm expm accuracy
    import scipy.linalg as sla
    n = 100; th = 2*np.pi*np.random.default_rng(0).random(n*(n-1)//2); K = o.construct_anti_symmetric(n, th)
    ch = expm_chain(K); print("s", ch["s"], "expm err", np.max(np.abs(ch["R"][-1]-sla.expm(K))), "orth", np.max(np.abs(ch["R"][-1].T@ch["R"][-1]-np.eye(n))))
    G = np.random.default_rng(1).standard_normal((n, n))
    Lt = dexpm_chain(ch, G.T); Lf = o._frechet_adjoint(K, G)
    print("frechet rel err", np.max(np.abs(Lt.T-Lf))/np.max(np.abs(Lf)))
