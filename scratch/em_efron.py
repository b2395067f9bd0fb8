"""Euler-Maclaurin evaluation of the Efron per-bin sums vs direct summation (longdouble)."""
import numpy as np, math, sys
Q = 16
B = [1/6, -1/30, 1/42, -1/30, 5/66]   # B2..B10
def psi_phi(z):
    if z < 0.03125:
        ps = 0.0; ph = 0.0
        for j in range(13, -1, -1):
            ps = ps * z + 1.0/(j+1); ph = ph * z + 1.0/(j+2)
        return ps, ph
    ps = -math.log1p(-z)/z
    return ps, (ps-1.0)/z
def em(m, r, K=5):
    a = r / m
    # L: largest integer with 1 - a L >= Q a  (L <= m)
    L = m if (1.0 - r) >= Q * a else int(math.floor(1.0/a - Q)) if a > 0 else m
    L = max(0, min(m, L))
    if L < 2: L = 0
    st = sg = sf = 0.0
    if L > 0:
        z = a * L
        ps, ph = psi_phi(z)
        u = 1.0/(1.0 - z)
        st = -L*z*(ps-ph) - 0.5*math.log1p(-z)
        sg = L*ps + 0.5*(1.0 - u)
        sf = (L*L/m)*ph - 0.5*(L/m)*u
        u2 = u*u
        a2 = a*a
        pw_odd = u      # (1-z)^-(2k-1)
        pw_even = u2    # (1-z)^-(2k)
        ak = a          # a^(2k-1)
        for k in range(1, K+1):
            b = B[k-1]
            st += -b/(2*k*(2*k-1)) * ak * (pw_odd - 1.0)
            sg += b/(2*k) * ak * (pw_even - 1.0)
            sf += b/(2*k) * (ak/a if a>0 else (1.0 if k==1 else 0.0)) / m * (pw_even - 1.0)
            pw_odd *= u2; pw_even *= u2; ak *= a2
    for l in range(L, m):
        x = 1.0 - l*a
        st += math.log(x); sg += 1.0/x; sf += (l/m)/x
    return st, sg, sf, m-L
def direct(m, r):
    l = np.arange(m, dtype=np.longdouble)
    x = 1 - l*(np.longdouble(r)/m)
    return float(np.log(x).sum()), float((1/x).sum()), float(((l/m)/x).sum())
rng = np.random.default_rng(0)
worst = [0,0,0]; wt=0
cases = []
for m in [1,2,3,5,16,17,18,19,33,100,1000,1250,5000,100000,3000000]:
    for r in [0.0,1e-12,1e-6,1e-3,0.01,0.1,0.3,0.5,0.9,0.99,0.999,1-1e-6,1-1e-9,1.0]:
        cases.append((m,r))
for _ in range(300):
    m = int(10**rng.uniform(0,5)); r = float(rng.choice([rng.uniform(0,1), 1-10**rng.uniform(-8,0), 10**rng.uniform(-8,0)]))
    cases.append((m, min(max(r,0.0),1.0)))
for m,r in cases:
    a = em(m,r); d = direct(m,r)
    errs = [abs(a[i]-d[i])/max(abs(d[i]),1e-300) if d[i]!=0 else abs(a[i]) for i in range(3)]
    # abs error relative to scale m for log-sum (sum can be near 0)
    errs[0] = abs(a[0]-d[0])/max(abs(d[0]), 1e-3*m*r+1e-30)
    wt=max(wt,a[3])
    for i in range(3):
        if errs[i] > worst[i]: worst[i]=errs[i]; print("new worst",i,m,r,errs[i],a[i],d[i])
print("worst rel errs", worst, "max tail", wt)
