import numpy as np
from scipy.special import erf, ndtr, log_ndtr
from scipy.optimize import least_squares
xs = np.linspace(0, 9, 20001)[1:]
phi = ndtr(xs)
logit = log_ndtr(xs) - log_ndtr(-xs)   # exact logit(Phi(x)), odd function
gelu = xs*phi
gelu_neg = -xs*ndtr(-xs)
def model(c, x):
    s = x*x
    q = np.zeros_like(x)
    for ck in c[::-1]:
        q = q*s + ck
    return q*x
def resid(c):
    p = model(c, xs)
    phat = 1/(1+np.exp(-p))
    # absolute error of GELU for +x and -x
    e1 = xs*phat - gelu
    e2 = -xs*(1-phat) - gelu_neg   # sigmoid(-p) = 1 - sigmoid(p)
    return np.concatenate([e1, e2])
for deg in (3,4,5,6):
    c0 = np.zeros(deg); c0[0]=1.5958; 
    if deg>1: c0[1]=0.0714
    r = least_squares(resid, c0, xtol=1e-15, ftol=1e-15, gtol=1e-15)
    print(deg, r.x, np.abs(resid(r.x)).max())
print("--- minimax-ish")
def fit_p(c0, power, xmax=9):
    def rp(c):
        r = resid(c)
        return np.sign(r)*np.abs(r/1e-6)**(power/2)
    r = least_squares(rp, c0, xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
    return r.x
c5 = np.array([1.59565838e+00, 7.29314157e-02, -2.45941020e-04, -6.19073477e-05, 2.28182475e-06])
for npar in (5, 6, 7):
    c = np.concatenate([c5, np.zeros(npar-5)])
    for power in (2, 4, 8, 16):
        with np.errstate(over='ignore'):
            c = fit_p(c, power)
            print(npar, power, np.abs(resid(c)).max())
    print(repr(c))
    # float32 check of the actual evaluation order
    c32 = c.astype(np.float32)
    x = np.linspace(-9, 9, 400001).astype(np.float32)
    s = x*x
    q = np.full_like(x, c32[-1])
    for ck in c32[-2::-1]:
        q = q*s + ck
    t = q*x
    y = x/(np.float32(1)+np.exp(-t))
    ref = x.astype(np.float64)*ndtr(x.astype(np.float64))
    print("f32 max abs err", np.abs(y-ref).max(), "at", x[np.abs(y-ref).argmax()])
