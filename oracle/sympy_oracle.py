"""Second, independent oracle: sympy restatement of the reference equations, evaluated with mpmath at
50 significant digits, Jacobians by SYMBOLIC differentiation (as CasADi's SX::jacobian does).

TEST INFRASTRUCTURE ONLY.  It exists to pin oracle/kite_oracle.hpp (dual-number C++) against an
independent restatement in a different language and differentiation method, and to generate the
committed golden vectors under tests/golden/ (see scripts/make_golden.py).

Follows /root/reference/src/kite_model/kite.cpp:197-322 (standard), :448-573 (identification variant),
:622-661 (rigid body); kite_math/kitemath.cpp:9-51; kite_math/pseudospectral/chebyshev.hpp:119-271;
kite_control/kiteNMPF.cpp:58-111; kite_estimation/kiteEKF.cpp:75-98.
"""
import mpmath as mp
import sympy as sp

mp.mp.dps = 50

ID_PARAM_NAMES = ["CL0", "CLa_total", "CD0_total", "CYb", "Cm0", "Cma", "Cnb", "Clb", "CLq", "Cmq", "CYr", "Cnr",
                  "Clr", "CYp", "Clp", "Cnp", "CLde", "CYdr", "Cmde", "Cndr", "Cldr"]


def qmul(a, b):
    """Hamilton product, scalar first (kitemath.cpp:9-23)."""
    s1, v1 = a[0], sp.Matrix(a[1:4])
    s2, v2 = b[0], sp.Matrix(b[1:4])
    s = s1 * s2 - v1.dot(v2)
    v = v1.cross(v2) + s1 * v2 + s2 * v1
    return [s, v[0], v[1], v[2]]


def qinv(a):
    return [a[0], -a[1], -a[2], -a[3]]


def heaviside(x, K):
    return K / (1 + sp.exp(-4 * x))


def R(x):
    """exact rational from the decimal literal / float repr (so 0.1 means 1/10 like the YAML text)."""
    return sp.Rational(repr(float(x))) if not isinstance(x, str) else sp.Rational(x)


def build_rhs(cfg, kind="kite"):
    """Returns (x syms[13], u syms[3], p syms[21 or 0], f exprs[13])."""
    x = sp.symbols("x0:13", real=True)
    u = sp.symbols("u0:3", real=True)
    v = x[0:3]; w = x[3:6]; r = x[6:9]; q = list(x[9:13])
    if kind == "rigid_body":
        vi = qmul(qmul(q, [0, *v]), qinv(q))[1:4]
        lam = -10
        qw = qmul(q, [0, *w])
        qq1 = sum(qi * qi for qi in q) - 1
        qdot = [sp.Rational(1, 2) * qw[i] + sp.Rational(1, 2) * lam * q[i] * qq1 for i in range(4)]
        return x, u, (), [0, 0, 0, 0, 0, 0, *vi, *qdot]

    g = R("9.80665"); ro = R("1.2985")
    geo, ine, aer, tet = cfg["geometry"], cfg["inertia"], cfg["aerodynamic"], cfg["tether"]
    b, c, AR, S = R(geo["b"]), R(geo["c"]), R(geo["AR"]), R(geo["S"])
    Mass, Ixx, Iyy, Izz, Ixz = (R(ine[k]) for k in ("mass", "Ixx", "Iyy", "Izz", "Ixz"))
    e_o, Cn0, Cl0 = R(aer["e_oswald"]), R(aer["Cn0"]), R(aer["Cl0"])
    Ks, Kd, Lt = R(tet["Ks"]), R(tet["Kd"]), R(tet["length"])
    rx, ry, rz = (R(tet.get(k, 0.0)) for k in ("rx", "ry", "rz"))
    if kind == "kite_id":
        p = sp.symbols("p0:21", real=True)
        co = dict(zip(ID_PARAM_NAMES, p))
        eps = 0
    else:
        p = ()
        co = {k: R(aer[k]) for k in ID_PARAM_NAMES}
        eps = R("1e-4")
    T_, dE, dR = u

    V2 = v[0] ** 2 + v[1] ** 2 + v[2] ** 2
    V = sp.sqrt(V2)
    ss = sp.asin(v[1] / (V + eps))
    aoa = sp.atan2(v[2], v[0] + eps)
    qd = sp.Rational(1, 2) * ro * V2
    CL = co["CL0"] + co["CLa_total"] * aoa
    CD = co["CD0_total"] + CL ** 2 / (sp.pi * e_o * AR)
    LIFT = CL * qd * S + (sp.Rational(1, 4) * co["CLq"] * c * S * ro) * V * w[1]
    DRAG = CD * qd * S
    SF = (co["CYb"] * ss + co["CYdr"] * dR) * qd * S + sp.Rational(1, 4) * (co["CYr"] * w[2] + co["CYp"] * w[0]) * (b * ro * S) * V
    q_aoa = [sp.cos(aoa / 2), 0, sp.sin(aoa / 2), 0]
    q_ss = [sp.cos(-ss / 2), 0, 0, sp.sin(-ss / 2)]
    qwb = qmul(q_aoa, q_ss)
    Faero = qmul(qmul(qinv(qwb), [0, -DRAG, 0, -LIFT]), qwb)[1:4]
    Zde = (-co["CLde"]) * dE * qd * S
    FdE = qmul(qmul(qinv(q_aoa), [0, 0, 0, Zde]), q_aoa)[1:4]
    Faero = [Faero[0] + FdE[0], Faero[1] + FdE[1] + SF, Faero[2] + FdE[2]]
    G_b = qmul(qmul(qinv(q), [0, 0, 0, g]), q)[1:4]
    T_b = [T_, 0, 0]
    d_ = sp.sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2)
    Rv = d_ - Lt
    Rs = [-Rv * (ri / d_) for ri in r]
    vi = qmul(qmul(q, [0, *v]), qinv(q))[1:4]
    rdv = r[0] * vi[0] + r[1] * vi[1] + r[2] * vi[2]
    Rd = [(-ri / d_) * rdv / d_ for ri in r]
    hv = heaviside(d_ - Lt, 1)
    Rt = [(Ks * Rs[i] + Kd * Rd[i]) * hv for i in range(3)]
    R_b = qmul(qmul(qinv(q), [0, *Rt]), q)[1:4]
    wxv = sp.Matrix(w).cross(sp.Matrix(v))
    v_dot = [(Faero[i] + T_b[i] + R_b[i]) / Mass + G_b[i] - wxv[i] for i in range(3)]
    L = (Cl0 + co["Clb"] * ss + co["Cldr"] * dR) * qd * S * b + (co["Clr"] * w[2] + co["Clp"] * w[0]) * (sp.Rational(1, 4) * ro * b ** 2 * S) * V
    M = (co["Cm0"] + co["Cma"] * aoa + co["Cmde"] * dE) * qd * S * c + co["Cmq"] * (sp.Rational(1, 4) * S * c ** 2 * ro) * w[1] * V
    N = (Cn0 + co["Cnb"] * ss + co["Cndr"] * dR) * qd * S * b + (co["Cnp"] * w[0] + co["Cnr"] * w[2]) * (sp.Rational(1, 4) * S * b ** 2 * ro) * V
    Maero = qmul(qmul(qinv(q_aoa), [0, L, M, N]), q_aoa)[1:4]
    Mt = sp.Matrix([rx, ry, rz]).cross(sp.Matrix(R_b))
    J = sp.Matrix([[Ixx, 0, Ixz], [0, Iyy, 0], [Ixz, 0, Izz]])
    wv = sp.Matrix(w)
    w_dot = J.inv() * (sp.Matrix(Maero) + Mt - wv.cross(J * wv))
    lam = -5
    qw = qmul(q, [0, *w])
    qq1 = sum(qi * qi for qi in q) - 1
    q_dot = [sp.Rational(1, 2) * qw[i] + sp.Rational(1, 2) * lam * q[i] * qq1 for i in range(4)]
    f = [*v_dot, *list(w_dot), *vi, *q_dot]
    build_rhs.last_aero = list(Faero)            # Function "Aero" (kite.cpp:330): body-frame aerodynamic force
    return x, u, p, f


class SymModel:
    """mpmath-evaluated f, df/dx, df/du for one model kind."""

    def __init__(self, cfg, kind="kite"):
        self.kind = kind
        x, u, p, f = build_rhs(cfg, kind)
        self.x, self.u, self.p = x, u, p
        fm = sp.Matrix(f)
        args = [*x, *u, *p]
        self._f = sp.lambdify(args, list(fm), modules="mpmath", cse=True)
        self._aero = sp.lambdify(args, list(getattr(build_rhs, "last_aero", [0, 0, 0])), modules="mpmath", cse=True) if kind != "rigid_body" else None
        Jx = fm.jacobian(sp.Matrix(x))
        Ju = fm.jacobian(sp.Matrix(u))
        self._J = sp.lambdify(args, [list(Jx), list(Ju)], modules="mpmath", cse=True)

    def f(self, x, u, p=()):
        a = [mp.mpf(t) for t in (*x, *u, *p)]
        return [mp.mpf(t) for t in self._f(*a)]

    def aero(self, x, u, p=()):
        a = [mp.mpf(t) for t in (*x, *u, *p)]
        return [mp.mpf(t) for t in self._aero(*a)]

    def jac(self, x, u, p=()):
        a = [mp.mpf(t) for t in (*x, *u, *p)]
        jx, ju = self._J(*a)
        Jx = [[mp.mpf(jx[i * 13 + j]) for j in range(13)] for i in range(13)]
        Ju = [[mp.mpf(ju[i * 3 + j]) for j in range(3)] for i in range(13)]
        return Jx, Ju

    def rk4_step(self, x, u, h, p=()):
        h = mp.mpf(h)
        x = [mp.mpf(t) for t in x]
        k1 = self.f(x, u, p)
        k2 = self.f([x[i] + h / 2 * k1[i] for i in range(13)], u, p)
        k3 = self.f([x[i] + h / 2 * k2[i] for i in range(13)], u, p)
        k4 = self.f([x[i] + h * k3[i] for i in range(13)], u, p)
        return [x[i] + (h / 6) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) for i in range(13)]

    def rk4_step_sens(self, x, u, h, p=()):
        """Phi = d x+/dx, Gamma = d x+/du by chaining symbolic stage Jacobians through the tableau."""
        h = mp.mpf(h)
        x = [mp.mpf(t) for t in x]
        n = 13
        E = mp.matrix(n, 16)
        for i in range(n):
            E[i, i] = 1
        B = lambda Ju: mp.matrix([[0] * 13 + [Ju[i][j] for j in range(3)] for i in range(n)])
        ks, Ss = [], []
        a = [mp.mpf(0), h / 2, h / 2, h]
        xi, dXi = x, E.copy()
        for st in range(4):
            if st > 0:
                xi = [x[i] + a[st] * ks[st - 1][i] for i in range(n)]
                dXi = E + a[st] * Ss[st - 1]
            k = self.f(xi, u, p)
            Jx, Ju = self.jac(xi, u, p)
            S = mp.matrix(Jx) * dXi + B(Ju)
            ks.append(k); Ss.append(S)
        xn = [x[i] + (h / 6) * (ks[0][i] + 2 * ks[1][i] + 2 * ks[2][i] + ks[3][i]) for i in range(n)]
        D = E + (h / 6) * (Ss[0] + 2 * Ss[1] + 2 * Ss[2] + Ss[3])
        Phi = [[D[i, j] for j in range(13)] for i in range(n)]
        Gam = [[D[i, 13 + j] for j in range(3)] for i in range(n)]
        return xn, Phi, Gam


# ---- Chebyshev operators (chebyshev.hpp:119-232), mpmath ---------------------------------------
def cheb_points(P):
    return [mp.cos(mp.mpf(k) * mp.pi / P) for k in range(P + 1)]


def cheb_diff(P):
    n = P + 1
    xs = cheb_points(P)
    c = [(-1) ** k * (2 if k in (0, P) else 1) for k in range(n)]
    Dn = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            Dn[i, j] = (mp.mpf(c[i]) / c[j]) / ((xs[i] - xs[j]) + (1 if i == j else 0))
    D = Dn.copy()
    for i in range(n):
        D[i, i] = Dn[i, i] - sum(Dn[i, j] for j in range(n))
    return D


def cheb_weights(P):
    """Clenshaw-Curtis (Trefethen clencurt)."""
    n = P + 1
    theta = [mp.mpf(k) * mp.pi / P for k in range(n)]
    w = [mp.mpf(0)] * n
    v = [mp.mpf(1)] * (P - 1)
    if P % 2 == 0:
        w[0] = w[P] = mp.mpf(1) / (P * P - 1)
        for k in range(1, P // 2):
            v = [v[i - 1] - 2 * mp.cos(2 * k * theta[i]) / (4 * k * k - 1) for i in range(1, P)]
        v = [v[i - 1] - mp.cos(P * theta[i]) / (P * P - 1) for i in range(1, P)]
    else:
        w[0] = w[P] = mp.mpf(1) / (P * P)
        for k in range(1, (P - 1) // 2 + 1):
            v = [v[i - 1] - 2 * mp.cos(2 * k * theta[i]) / (4 * k * k - 1) for i in range(1, P)]
    for i in range(1, P):
        w[i] = 2 * v[i - 1] / P
    return w


def cheb_compdiff(P, S):
    m, n = S * P + 1, P + 1
    D = cheb_diff(P)
    if S < 2:
        return D
    Cm = mp.matrix(m, m)
    for i in range(n):
        for j in range(n):
            Cm[m - n + i, m - n + j] = D[i, j]
    for k in range(0, (S - 1) * P, P):
        for i in range(P):
            for j in range(n):
                Cm[k + i, k + j] = D[i, j]
    return Cm


def colloc_eval(model, z, P, S, t0, tf, sx, su):
    """G and the node Jacobian blocks of the scaled augmented dynamics (kiteNMPF.cpp:58-111)."""
    M = S * P + 1
    tau = (mp.mpf(tf) - mp.mpf(t0)) / (2 * S)
    Cm = cheb_compdiff(P, S)
    X = [[mp.mpf(z[k * 15 + i]) for i in range(15)] for k in range(M)]
    U = [[mp.mpf(z[M * 15 + k * 4 + i]) for i in range(4)] for k in range(M)]
    sx = [mp.mpf(t) for t in sx]; su = [mp.mpf(t) for t in su]
    F, JX, JU = [], [], []
    for k in range(M):
        x = [X[k][i] / sx[i] for i in range(15)]
        u = [U[k][i] / su[i] for i in range(4)]
        f = model.f(x[:13], u[:3])
        Jx, Ju = model.jac(x[:13], u[:3])
        fa = [*f, x[14], u[3]]
        F.append([sx[i] * fa[i] for i in range(15)])
        jx = [[mp.mpf(0)] * 15 for _ in range(15)]
        ju = [[mp.mpf(0)] * 4 for _ in range(15)]
        for i in range(13):
            for j in range(13):
                jx[i][j] = sx[i] * Jx[i][j] / sx[j]
            for j in range(3):
                ju[i][j] = sx[i] * Ju[i][j] / su[j]
        jx[13][14] = sx[13] / sx[14]
        ju[14][3] = sx[14] / su[3]
        JX.append(jx); JU.append(ju)
    G = []
    for k in range(M):
        for i in range(15):
            G.append(sum(Cm[k, l] * X[l][i] for l in range(M)) - tau * F[k][i])
    return G, JX, JU


def ekf_predict(model, x, u, dt, Pc, W):
    xn = model.rk4_step(x, u, dt)
    Jx, _ = model.jac(x, u)
    A = mp.matrix(Jx) * mp.mpf(dt) + mp.eye(13)
    Pn = A * mp.matrix(Pc) * A.T + mp.matrix(W)
    return xn, Pn
