"""Python big-int model of the PRODUCT's decoder (mpc-protocols_b200/csrc/robust.cuh), NOT of the reference's -- test infrastructure.

The reference decodes with Gao's algorithm inside its OEC rounds (robust_interpolate.rs:94-157, 456-538, 579-628; restated in
oracle/pymodel.py).  The CUDA path reaches the same (coefficients, path, flags) with a different algorithm: weighted syndromes ->
inversion-free Berlekamp-Massey -> Chien search -> Forney error values, one "fast" attempt over all supplied shares and, only when that
fails, the literal round-by-round attempts (robust.cuh:1-30 argues why the outcomes must coincide).  This file restates that
algorithm with the same recurrences, thresholds and path rule as the kernels, so that tests/test_syndrome_equivalence.py can check
the equivalence claim against the reference-faithful model on the CPU (`-m "not gpu"`), independently of the GPU parity tests that
compare the kernels themselves with the oracle.  Nothing of the product imports it.  Parity is unpinned at the arkworks byte level like
the rest of oracle/ (no Rust toolchain here)."""
from __future__ import annotations

from . import pymodel as pm

R = pm.R_MOD
inv = pm.inv


def _weights(xs):
    """u_i = 1 / prod_{l != i} (x_i - x_l)   (RobustArgs::u2 / uinv)"""
    out = []
    for i, xi in enumerate(xs):
        p = 1
        for l, xl in enumerate(xs):
            if l != i:
                p = p * (xi - xl) % R
        out.append(inv(p))
    return out


def attempt(xs, ys, d: int, max_l: int, values: str = "forney"):
    """rs_attempt (robust.cuh:189-446) on the points xs (distinct, id-sorted prefix) with received values ys.
    Returns [(position, error value)] (ascending position) or None when the attempt fails.
    values = "forney": Omega = S*Lambda mod z^L, as the kernels compute today.
    values = "hk": the same error values WITHOUT Omega (Horiguchi-Koetter form), from the auxiliary polynomial B
    that Berlekamp-Massey keeps anyway.  With j* = nsyn - shift the iteration of the last length change, B its locator before that change and
    bdis the discrepancy it met there, Lambda*(S*B mod z^deg) - B*Omega = -bdis * Lambda(0) * z^j* exactly (both sides have degree <= L + L_B - 1 = j*
    and S*Lambda = Omega as power series), so at a root z of Lambda:  Omega(z) = bdis * Lambda(0) * z^j* / B(z)."""
    P = len(xs)
    nsyn = P - (d + 1)
    u = _weights(xs)
    # syndromes S_j = sum_i u_i x_i^j y_i: zero for every word that lies on a polynomial of degree <= d
    syn = [sum(u[i] * pow(xs[i], j, R) * ys[i] for i in range(P)) % R for j in range(nsyn)]
    # inversion-free Berlekamp-Massey: Lambda <- bdis*Lambda - delta * z^shift * B
    lam, bp = [1], [1]
    bdis, L, shift = 1, 0, 1
    for j in range(nsyn):
        delta = sum(lam[l] * syn[j - l] for l in range(min(L, j) + 1)) % R
        if delta == 0:
            shift += 1
            continue
        grow = 2 * L <= j
        new_l = j + 1 - L if grow else L
        if new_l > max_l:
            return None
        new = []
        for l in range(new_l + 1):
            a = lam[l] if l < len(lam) else 0
            bi = l - shift
            b = bp[bi] if 0 <= bi < len(bp) else 0
            new.append((a * bdis - b * delta) % R)
        if grow:
            bp = lam[: L + 1] + [0] * (L + 1 - len(lam))
            bdis, L, shift = delta, new_l, 1
        else:
            shift += 1
        lam = new
    if L == 0:
        return []
    lam = lam + [0] * (L + 1 - len(lam))
    omega = [sum(lam[k] * syn[l - k] for k in range(l + 1)) % R for l in range(L)]   # S*Lambda mod z^L
    roots = [i for i in range(P) if pm.p_eval(lam[: L + 1], inv(xs[i])) == 0]           # Chien search over the prefix
    if len(roots) != L:
        return None
    dlam = [(l * lam[l]) % R for l in range(1, L + 1)]
    out = []
    for i in roots:
        z = inv(xs[i])
        den = pm.p_eval(dlam, z)
        if den == 0:
            return None
        if values == "hk":
            bz = pm.p_eval(bp, z)
            if bz == 0:
                return None
            om = bdis * lam[0] % R * pow(z, nsyn - shift, R) % R * inv(bz) % R
        else:
            om = pm.p_eval(omega, z)
        c = (-xs[i] * om) % R * inv(den) % R                        # Forney
        out.append((i, c * inv(u[i]) % R))                          # e_i = c_i * prod_{l != i}(x_i - x_l)
    return out


def robust_recover_secret(shares, n: int, t: int):
    """The product's route for one codeword (hbmpc_robust_interpolate_batch): optimistic check, fast attempt on all supplied shares
    with the path rule of robust_kernel, then the literal rounds.  shares = [(id, value, degree)] in arrival order.
    Returns dict(coeffs (trimmed), secret, path, flags) or None (DecodingError), like oracle.pymodel.robust_recover_secret."""
    d = shares[0][2]
    srt = sorted(((s[0], s[1]) for s in shares), key=lambda s: s[0])
    S, needed = len(srt), d + t + 1
    xs = [pm.domain_element(n, i) for i, _ in srt]
    ys = [y for _, y in srt]

    def finish(poly, path):
        poly = pm.p_norm(poly)
        return {"coeffs": poly, "secret": pm.p_eval(poly, 0), "path": path,
                "flags": [pm.p_eval(poly, pm.domain_element(n, s[0])) != s[1] for s in shares]}

    # optimistic: the examined prefix of d+t+1 shares lies on one polynomial of degree <= d
    base = pm.lagrange_interpolate(xs[: d + 1], ys[: d + 1])
    if all(pm.p_eval(base, xs[i]) == ys[i] for i in range(needed)):
        return finish(base, 0)
    rmax = min(t, S - needed)

    def corrected(P, errs):
        y2 = list(ys[:P])
        for i, e in errs:
            y2[i] = (y2[i] - e) % R
        return pm.lagrange_interpolate(xs[: d + 1], y2[: d + 1])

    # fast attempt: ALL supplied shares, radius min(t, floor((S-d-1)/2))
    errs = attempt(xs, ys, d, min(t, (S - d - 1) // 2))
    if errs is not None:
        q = 0
        for r in range(1, rmax + 1):   # the reference accepts in the first round whose prefix holds at most r of the errors
            while q < len(errs) and errs[q][0] < needed + r:
                q += 1
            if q <= r:
                return finish(corrected(S, errs), r)
        return None
    for r in range(1, rmax + 1):       # exact path: the literal OEC loop, radius r on the prefix of d+t+1+r shares
        P = needed + r
        errs = attempt(xs[:P], ys[:P], d, r)
        if errs is not None:
            return finish(corrected(P, errs), r)
    return None
