"""Oracle: RAMBO-on-diet flat phase space 2 -> n with pT / deltaR / rapidity cuts — CPU, float64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the ``pdf_active=False`` path of
  /root/reference/nisrep/PhaseSpace/flat_phase_space_generator.py:81-113,139-441
  /root/reference/nisrep/PhaseSpace/utils.py:5-81,134-187
as straight-line tensor code (one formula per reference step, cited inline).  The root of the
mass-generation polynomial is found with the reference's own integer-lattice bisection (:313-359) so
that the oracle tracks the reference to rounding; the CUDA kernel uses a Newton iteration instead and
is compared with a tolerance.
"""
import math

import torch

EPS_SQRT = float(torch.finfo(torch.float64).eps) ** 0.5     # utils.py:151 eps=np.finfo(float).eps**0.5
HUGE = float(torch.finfo(torch.float64).max)                 # utils.py:151 huge


class PhaseSpaceGeneratorError(Exception):
    pass


def n_dim_phase_space(n_final):
    """flat_phase_space_generator.py:48-54."""
    return 0 if n_final == 1 else 3 * n_final - 4


def flat_weight(E_cm, n):
    """flat_phase_space_generator.py:81-97 — massless n-body volume."""
    if n == 1:
        return 1.0
    return math.pow(2 * math.pi, 4 - 3 * n) * math.pow(math.pi / 2.0, n - 1) * \
        (math.pow(E_cm ** 2, n - 2) / (math.factorial(n - 1) * math.factorial(n - 2)))


def rho(M, N, m):
    """flat_phase_space_generator.py:107-113."""
    Msqr = M ** 2
    return ((Msqr - (N + m) ** 2) * (Msqr - (N - m) ** 2)) ** 0.5 / (8.0 * Msqr)


def massless_map(x, e):
    """flat_phase_space_generator.py:101-103."""
    return (x ** e) * ((e + 1) - e * x)


def bisect(v, n_final, target=1e-16, max_level=600):
    """flat_phase_space_generator.py:313-359, verbatim control flow (batch-max stopping rule)."""
    if v.shape[1] == 0:
        return v
    e = torch.arange(n_final - 2, 0, step=-1, dtype=torch.float64).unsqueeze(0).repeat(v.shape[0], 1)
    level = 0
    left = torch.zeros_like(v)
    right = torch.ones_like(v)
    check = -torch.ones_like(v)
    u = -torch.ones_like(v)
    error = torch.ones_like(v)
    step = max_level / 10
    ml = step
    old = 100
    while torch.max(error) > target and ml < 10 * step:
        while level < ml:
            u = (left + right) * (0.5 ** (level + 1))
            check = massless_map(u, e)
            left = left * 2.0
            right = right * 2.0
            adder = torch.where(v <= check, torch.full_like(v, -0.5), torch.full_like(v, 0.5))
            left = left + (adder + 0.5)
            right = right + (adder - 0.5)
            level += 1
        error = torch.abs(1.0 - check / v)
        ml = ml + step
        new = torch.max(error)
        if new >= old:
            break
        old = new
    return u


def _pseudorap(p):
    """utils.py:151-166.  p: [..., 4] -> [...]."""
    pt = torch.sqrt(p[..., 1] ** 2 + p[..., 2] ** 2)
    th = torch.atan2(pt, p[..., 3])
    eta = -torch.log(torch.tan(th / 2.0))
    return torch.where((pt < EPS_SQRT) & (p[..., 3].abs() < EPS_SQRT), torch.full_like(eta, HUGE), eta)


def _delphi(p1, p2):
    """utils.py:170-180."""
    pt1 = torch.sqrt(p1[:, 1] ** 2 + p1[:, 2] ** 2)
    pt2 = torch.sqrt(p2[:, 1] ** 2 + p2[:, 2] ** 2)
    tmp = (p1[:, 1] * p2[:, 1] + p1[:, 2] * p2[:, 2]) / (pt1 * pt2)
    r = torch.where(tmp.abs() > 1, torch.acos(tmp / tmp.abs()), torch.acos(tmp))
    return torch.where((pt1 == 0.0) | (pt2 == 0.0), torch.full_like(r, HUGE), r)


def delta_r(p1, p2):
    """utils.py:182-187."""
    return torch.sqrt((_pseudorap(p1) - _pseudorap(p2)) ** 2 + _delphi(p1, p2) ** 2)


def pdf_density(pdf, pdg, x, scale2):
    """get_pdfQ2, flat_phase_space_generator.py:120-137: xf(x, Q^2) / x for gluons and quarks, 1 otherwise."""
    if pdf is None:
        return torch.ones_like(x)
    if pdg not in [21] and abs(pdg) not in range(1, 7):
        return torch.ones_like(x)
    f = pdf.xfxQ2(pdg, x, scale2)
    return torch.tensor(f, dtype=torch.float64) / x


def lab_boost(momenta, xb_1, xb_2):
    """boost_to_lab_frame, utils.py:134-146 with boost_tt :83-106.  The reference decides batch-wide: if ANY event has
    a reference vector x1 p1 + x2 p2 at rest (x1 == x2 for equal beams) nothing is boosted, otherwise every event
    is (its torch.where sees the tensor boost_tt already changed in place)."""
    ref = momenta[:, 0, :] * xb_1.unsqueeze(-1) + momenta[:, 1, :] * xb_2.unsqueeze(-1)
    if bool(((ref[:, 1:] ** 2).sum(-1) == 0).any()):
        return momenta
    bv = (ref[:, 1:] / ref[:, 0:1]).unsqueeze(1)                    # boostVector_t, [B,1,3]
    b2 = (bv * bv).sum(-1)                                           # [B,1]
    gamma = 1.0 / torch.sqrt(1.0 - b2)
    bp = (momenta[:, :, 1:] * bv).sum(-1)
    gamma2 = torch.where(b2 > 0, (gamma - 1.0) / b2, torch.zeros_like(b2))
    factor = gamma2 * bp + gamma * momenta[:, :, 0]
    space = momenta[:, :, 1:] + factor.unsqueeze(-1) * bv
    e = gamma * (momenta[:, :, 0] + bp)
    return torch.cat((e.unsqueeze(-1), space), -1)


def generate_kinematics(E_cm, r, initial_masses, final_masses,
                        pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1, return_parts=False,
                        pdf=None, pdf_active=False, tau=True, pdgs=(0, 0), absolute_Ecm_min=1.0):
    """flat_phase_space_generator.py:139-308.

    r: [B, 3n-4] float64 (+ 2 columns when ``pdf_active``: tau / y_cm, or x2 / x1 with ``tau=False``).  Returns
    (momenta [B, 2+n, 4] (E,px,py,pz; CM frame), weight [B]); with ``return_parts`` also (flat*massive weight
    before cuts, cut factor in {0,1})."""
    r = torch.as_tensor(r, dtype=torch.float64)
    n = len(final_masses)
    if len(initial_masses) != 2:                                    # :76-79
        raise PhaseSpaceGeneratorError("only 2 incoming particles")
    if torch.isnan(r).any():                                        # :147-149
        raise PhaseSpaceGeneratorError("NaN random variables")
    m = torch.tensor(final_masses, dtype=torch.float64)
    collider_energy = E_cm
    B = r.shape[0]
    wgt_jac = torch.ones(B, dtype=torch.float64)
    xb_1 = torch.ones(B, dtype=torch.float64)
    xb_2 = torch.ones(B, dtype=torch.float64)
    if pdf_active:                                                  # :157-187
        full = r
        r = full[:, :-2]
        if tau:
            tau_min = (max(float(m.sum()), absolute_Ecm_min) / E_cm) ** 2
            tau_v = tau_min + (1.0 - tau_min) * full[:, -2]         # uniform_distr utils.py:124-132
            ycm_min = 0.5 * torch.log(tau_v)
            ycm = ycm_min + (-ycm_min - ycm_min) * full[:, -1]
            xb_1 = torch.sqrt(tau_v) * torch.exp(ycm)
            xb_2 = torch.sqrt(tau_v) * torch.exp(-ycm)
            E_cm = torch.sqrt(tau_v) * E_cm
            wgt_jac = wgt_jac * ((1.0 - tau_min) * (-ycm_min - ycm_min))
        else:
            xb_1 = full[:, -1]
            xb_2 = full[:, -2]
            E_cm = torch.sqrt(xb_1 * xb_2) * E_cm
        q2 = torch.ones_like(xb_1) * 91.188 ** 2
        x_cut = torch.where(xb_1 < 1e-4, torch.zeros_like(xb_1), torch.ones_like(xb_1))
        x_cut = torch.where(xb_2 < 1e-4, torch.zeros_like(x_cut), x_cut)
        wgt_jac = wgt_jac * (pdf_density(pdf, pdgs[0], xb_1, q2) * pdf_density(pdf, pdgs[1], xb_2, q2) * x_cut)
    assert r.shape[1] == n_dim_phase_space(n)                       # :191

    # (1) massless intermediate masses K_j, :204-210,:384,:363-370
    K = torch.zeros(B, n - 1, dtype=torch.float64)
    K[:, 0] = E_cm - m.sum()
    u = bisect(r[:, :n - 2], n)
    for i in range(2, n):
        K[:, i - 1] = torch.sqrt(u[:, i - 2] * K[:, i - 2] ** 2)
    # (2) flat weight, :372
    if torch.is_tensor(E_cm):                                       # :95-97
        w = math.pow(2 * math.pi, 4 - 3 * n) * math.pow(math.pi / 2.0, n - 1) * \
            (torch.pow(E_cm ** 2, n - 2) / (math.factorial(n - 1) * math.factorial(n - 2)))
    else:
        w = torch.full((B,), flat_weight(E_cm, n), dtype=torch.float64)
    w = w * wgt_jac
    # (3) massive intermediates and reweighting, :389-403
    msum = torch.flip(torch.cumsum(torch.flip(m, (-1,)), -1), (-1,))
    M = K + msum[:-1]
    w = w * 8.0 * rho(M[:, n - 2], m[n - 1], m[n - 2])
    if n > 2:
        w = w * torch.prod(rho(M[:, :n - 2], M[:, 1:], m[:n - 2]) / rho(K[:, :n - 2], K[:, 1:], 0.0)
                           * (M[:, 1:n - 1] / K[:, 1:n - 1]), -1)
    w = w * torch.pow(K[:, 0] / M[:, 0], 2 * n - 4)
    # (4) decay momenta, :226-228
    Mx = torch.cat((M, m[-1:].unsqueeze(0).repeat(B, 1)), -1)       # M_{n-1} = m_{n-1}
    q = 4.0 * Mx[:, :-1] * rho(Mx[:, :-1], Mx[:, 1:], m[:-1])
    # (5) angles, :230-243
    rnd = r[:, n - 2:3 * n - 4]
    ct = 2.0 * rnd[:, 0::2] - 1.0
    st = torch.sqrt(1.0 - ct ** 2)
    phi = 2 * math.pi * rnd[:, 1::2]
    cp = torch.cos(phi)
    sp0 = torch.sqrt(1.0 - cp ** 2)
    sp = torch.where(phi > math.pi, -sp0, sp0)
    # (6) sequential two-body decays, :250-278
    out = torch.zeros(B, 2 + n, 4, dtype=torch.float64)
    Q = torch.zeros(B, 4, dtype=torch.float64)
    Q[:, 0] = Mx[:, 0]
    for j in range(n - 1):
        p = torch.stack((torch.zeros(B, dtype=torch.float64),
                         q[:, j] * st[:, j] * cp[:, j], q[:, j] * st[:, j] * sp[:, j],
                         q[:, j] * ct[:, j]), -1)
        p[:, 0] = torch.sqrt((p[:, 1:] ** 2).sum(-1) + m[j] ** 2)   # set_square_t utils.py:5-19
        bv = Q[:, 1:] / Q[:, 0:1]                                    # boostVector_t utils.py:31-36
        b2 = (bv * bv).sum(-1)                                       # boost_t utils.py:58-81
        gamma = 1.0 / torch.sqrt(1.0 - b2)
        bp = (p[:, 1:] * bv).sum(-1)
        gamma2 = torch.where(b2 > 0, (gamma - 1.0) / b2, torch.zeros_like(b2))
        fac = gamma2 * bp + gamma * p[:, 0]
        ps = p[:, 1:] + fac.unsqueeze(1) * bv
        p = torch.cat((torch.sqrt((ps ** 2).sum(-1) + m[j] ** 2).unsqueeze(1), ps), -1)   # :265
        out[:, 2 + j] = p
        Qs = Q[:, 1:] - p[:, 1:]
        Q = torch.cat((torch.sqrt((Qs ** 2).sum(-1) + Mx[:, j + 1] ** 2).unsqueeze(1), Qs), -1)  # :271-275
    out[:, -1] = Q                                                   # :278
    # (7) beams, :408-441 (per event when E_cm is a tensor)
    m1, m2 = float(initial_masses[0]), float(initial_masses[1])
    Ev = E_cm if torch.is_tensor(E_cm) else torch.full((B,), float(E_cm), dtype=torch.float64)
    zero = torch.zeros_like(Ev)
    if m1 == 0.0 or m2 == 0.0:
        out[:, 0] = torch.stack((Ev / 2.0, zero, zero, Ev / 2.0), -1)
        out[:, 1] = torch.stack((Ev / 2.0, zero, zero, -Ev / 2.0), -1)
    else:
        M1, M2 = m1 ** 2, m2 ** 2
        E1 = (Ev ** 2 + M1 - M2) / Ev
        E2 = (Ev ** 2 - M1 + M2) / Ev
        Z = torch.sqrt(Ev ** 4 - 2 * Ev ** 2 * M1 - 2 * Ev ** 2 * M2 + M1 ** 2 - 2 * M1 * M2 + M2 ** 2) / Ev
        out[:, 0] = torch.stack((E1 / 2.0, zero, zero, Z / 2.0), -1)
        out[:, 1] = torch.stack((E2 / 2.0, zero, zero, -Z / 2.0), -1)
    cm = out                                                         # returned: the CM-frame clone (:282,:308)
    if pdf_active:
        out = lab_boost(out, xb_1, xb_2)                             # :283
    # (8) cuts on the (lab-frame) final state; pdf inactive: x1 = x2 = 1 and boost_to_lab_frame is the identity
    fin = out[:, 2:]
    ptmin = torch.sqrt(fin[:, :, 1] ** 2 + fin[:, :, 2] ** 2).abs().min(1).values
    cut = torch.where(ptmin < pT_mincut, torch.zeros_like(w), torch.ones_like(w))     # :285-288
    for i in range(n):
        for j in range(n):
            if i > j:                                                                  # :290-296
                cut = cut * torch.where(delta_r(fin[:, i], fin[:, j]).abs() < delR_mincut,
                                        torch.zeros_like(w), torch.ones_like(w))
    if rap_maxcut > 0:                                                                 # :298-301
        cut = cut * torch.where(rap_maxcut < _pseudorap(fin).max(1).values.abs(),
                                torch.zeros_like(w), torch.ones_like(w))
    # (9) :304-308
    weight = w * cut / (2.0 * (xb_1 * xb_2 * collider_energy ** 2))   # shat = x1 x2 s
    if return_parts:
        return cm, weight, w, cut
    return cm, weight


# ----------------------------------------------------------------------------------------------
# inverse map.  The reference has none (README.md:68-69: to do); this is the algebraic inverse of
# generate_kinematics above, pinned by round trips through the reference's own golden momenta (tests/test_oracle_golden.py).
# ----------------------------------------------------------------------------------------------
def invert_kinematics(E_cm, momenta, initial_masses, final_masses):
    """momenta [B, 2+n, 4] (E,px,py,pz; CM frame, beams first) -> (r [B, 3n-4], weight [B] without cuts).

    Written independently of the forward's arrangement: the parent system Q_j = sum_{i>=j} p_i, its mass from the
    Minkowski square, the daughter brought to the parent's rest frame by the textbook boost with beta = -Q/Q0,
    gamma = Q0/M, and the mass uniforms from the polynomial the forward inverts (massless_map, :101-105)."""
    mom = torch.as_tensor(momenta, dtype=torch.float64)
    n = len(final_masses)
    B = mom.shape[0]
    m = torch.tensor(final_masses, dtype=torch.float64)
    fin = mom[:, 2:, :]
    r = torch.zeros(B, 3 * n - 4, dtype=torch.float64)
    msum = torch.flip(torch.cumsum(torch.flip(m, (-1,)), -1), (-1,))           # msum[j] = sum_{i>=j} m_i
    Q = torch.zeros(B, 4, dtype=torch.float64)
    Q[:, 0] = E_cm                                                  # the sum of the final state, exactly (E_cm, 0, 0, 0)
    K = torch.full((B,), float(E_cm) - float(m.sum()), dtype=torch.float64)
    for j in range(n - 1):
        p = fin[:, j, :]
        M = torch.sqrt((Q[:, 0] ** 2 - (Q[:, 1:] ** 2).sum(-1)).clamp_min(0.0))
        beta = -Q[:, 1:] / Q[:, :1]
        gamma = (Q[:, 0] / M).unsqueeze(-1)
        bp = (beta * p[:, 1:]).sum(-1, keepdim=True)
        b2 = (beta ** 2).sum(-1, keepdim=True)
        # p' = p + [(gamma - 1) (beta.p) / beta^2 + gamma E] beta   (boost of the frame by velocity -beta ... applied with beta = -Q/Q0)
        fac = torch.where(b2 > 0, (gamma - 1.0) * bp / b2.clamp_min(1e-300), torch.zeros_like(bp)) + gamma * p[:, :1]
        prest = p[:, 1:] + fac * beta
        pm = torch.sqrt((prest ** 2).sum(-1))
        ct = torch.where(pm > 0, prest[:, 2] / pm.clamp_min(1e-300), torch.ones_like(pm)).clamp(-1.0, 1.0)
        phi = torch.atan2(prest[:, 1], prest[:, 0]) / (2.0 * math.pi)
        phi = torch.where(phi < 0, phi + 1.0, phi)
        r[:, n - 2 + 2 * j] = 0.5 * (ct + 1.0)
        r[:, n - 1 + 2 * j] = phi
        if j < n - 2:
            Q = Q - p
            Mn = torch.sqrt((Q[:, 0] ** 2 - (Q[:, 1:] ** 2).sum(-1)).clamp_min(0.0))
            Kn = Mn - msum[j + 1]
            u = ((Kn / K) ** 2).clamp(0.0, 1.0)
            e = n - 2 - j
            r[:, j] = (e + 1) * u ** e - e * u ** (e + 1)          # massless_map, :101-105
            K = Kn
    _, w = generate_kinematics(E_cm, r, initial_masses, final_masses)
    return r, w

