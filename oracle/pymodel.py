"""Independent Python big-int model of the reference hot path (TEST INFRASTRUCTURE ONLY).

This file is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it.  It is a *second*, independent restatement (plain Python
ints, `pow`, `%`) used to pin the C oracle (oracle/hbmpc_oracle.c) and to generate the
golden fixtures under tests/golden/ (tests/golden/make_golden.py).

PARITY UNPINNED at the arkworks byte level: the reference is Rust (no rustc/cargo in this
environment) and its field/polynomial arithmetic lives in ark-ff/ark-poly 0.5.0, which are not
vendored under /root/reference.  Canonical residues mod r are unique, so the pins that exist are
the reference's own known-answer tests, reproduced in
tests/test_oracle_kats.py.

Every function cites the reference lines it restates (paths relative to /root/reference/mpc/src).
Polynomials are Python lists of ints, lowest degree first, kept *normalised* (no trailing zero
coefficients; the zero polynomial is []), mirroring ark_poly::DensePolynomial.
"""
from __future__ import annotations

# ---------------------------------------------------------------------------------------------
# Field: ark_bls12_381::Fr  (ark-bls12-381 0.5.0; modulus, generator 7, two-adicity 32)
# ---------------------------------------------------------------------------------------------
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
GENERATOR = 7
TWO_ADICITY = 32
ROOT_2_32 = pow(GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
assert ROOT_2_32 == 10238227357739495823651030575849232062558860180284477541189508159991286009131
MONT_R = (1 << 256) % R_MOD
MONT_R2 = (MONT_R * MONT_R) % R_MOD

# ShareErrorCode, ffi/c_bindings/share/mod.rs:18-37
SHARE_SUCCESS = 0
INSUFFICIENT_SHARES = 1
DEGREE_MISMATCH = 2
ID_MISMATCH = 3
INVALID_INPUT = 4
TYPE_MISMATCH = 5
NO_SUITABLE_DOMAIN = 6
POLYNOMIAL_OPERATION_ERROR = 7
DECODING_ERROR = 8


class ShareErr(Exception):
    def __init__(self, code, msg=""):
        super().__init__(f"ShareErrorCode {code}: {msg}")
        self.code = code


def inv(a: int) -> int:
    return pow(a, R_MOD - 2, R_MOD)


# ---------------------------------------------------------------------------------------------
# Evaluation domain: common/mod.rs:51-68 -> GeneralEvaluationDomain::new(n) (ark-poly 0.5.0):
# radix-2 domain of size N = next_power_of_two(n); element(j) = w_N^j.
# ---------------------------------------------------------------------------------------------
def domain_size(n: int) -> int:
    if n <= 0:
        raise ShareErr(NO_SUITABLE_DOMAIN)
    N = 1
    while N < n:
        N <<= 1
    if N > (1 << TWO_ADICITY):
        raise ShareErr(NO_SUITABLE_DOMAIN)
    return N


def domain_gen(n: int) -> int:
    N = domain_size(n)
    return pow(ROOT_2_32, (1 << TWO_ADICITY) // N, R_MOD)


def domain_element(n: int, j: int) -> int:
    return pow(domain_gen(n), j, R_MOD)


# ---------------------------------------------------------------------------------------------
# DensePolynomial helpers (ark-poly 0.5.0 semantics)
# ---------------------------------------------------------------------------------------------
def p_norm(a):
    a = list(a)
    while a and a[-1] % R_MOD == 0:
        a.pop()
    return [x % R_MOD for x in a]


def p_degree(a):  # DensePolynomial::degree(): zero polynomial -> 0
    return 0 if not a else len(a) - 1


def p_add(a, b):
    n = max(len(a), len(b))
    return p_norm([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % R_MOD for i in range(n)])


def p_sub(a, b):
    n = max(len(a), len(b))
    return p_norm([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % R_MOD for i in range(n)])


def p_mul(a, b):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % R_MOD
    return p_norm(out)


def p_scale(a, s):
    return p_norm([(x * s) % R_MOD for x in a])


def p_eval(a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % R_MOD
    return acc


def p_divmod(a, b):
    """DenseOrSparsePolynomial::divide_with_q_and_r; b == 0 panics in the reference."""
    if not b:
        raise ZeroDivisionError("division by zero polynomial")
    if not a:
        return [], []
    if len(a) < len(b):
        return [], list(a)
    q = [0] * (len(a) - len(b) + 1)
    r = list(a)
    lead_inv = inv(b[-1])
    while r and len(r) >= len(b):
        cq = (r[-1] * lead_inv) % R_MOD
        sh = len(r) - len(b)
        q[sh] = cq
        for i, y in enumerate(b):
            r[sh + i] = (r[sh + i] - cq * y) % R_MOD
        r = p_norm(r)
    return p_norm(q), r


# ---------------------------------------------------------------------------------------------
# compute_shares: robust_interpolate.rs:52-82 == shamir.rs:158-196 (identical math).
# The reference draws a random degree-d polynomial, overwrites coeff 0 with the secret and takes the
# first n values of domain.fft(poly).  Here the coefficient vector is the input (no hidden RNG).
# ---------------------------------------------------------------------------------------------
def compute_shares(coeffs, n: int, degree: int):
    if n <= degree:
        raise ShareErr(INVALID_INPUT, "n must be greater than degree")
    if len(coeffs) != degree + 1:
        raise ShareErr(INVALID_INPUT, "need degree+1 coefficients")
    w = domain_gen(n)
    return [p_eval(p_norm(coeffs), pow(w, j, R_MOD)) for j in range(n)]


# make_vandermonde: common/share/mod.rs:31-45   V[j][k] = element(j)^k, j<n, k<=t
def make_vandermonde(n: int, t: int):
    w = domain_gen(n)
    return [[pow(w, j * k, R_MOD) for k in range(t + 1)] for j in range(n)]


# apply_vandermonde: common/share/mod.rs:50-76 (values only; id/degree checks live in the callers)
def apply_vandermonde(V, values):
    for row in V:
        if len(row) != len(values):
            raise ShareErr(INVALID_INPUT)
    return [sum(a * b for a, b in zip(row, values)) % R_MOD for row in V]


# lagrange_interpolate: common/mod.rs:134-165 (naive coefficient-form Lagrange)
def lagrange_interpolate(xs, ys):
    if len(xs) != len(ys):
        raise ShareErr(INVALID_INPUT)
    if len(set(xs)) != len(xs):
        raise ShareErr(INVALID_INPUT)
    result = []
    for j in range(len(xs)):
        num = [1]
        den = 1
        for m in range(len(xs)):
            if m != j:
                num = p_mul(num, [(-xs[m]) % R_MOD, 1])
                den = (den * (xs[j] - xs[m])) % R_MOD
        term = p_mul(num, p_norm([(ys[j] * inv(den)) % R_MOD]))
        result = p_add(result, term)
    return result


# NonRobustShare::recover_secret: shamir.rs:199-239.  shares = [(id, value)], all of degree `deg`.
def nonrobust_recover_secret(shares, n: int, deg: int):
    if not shares:
        raise ShareErr(INVALID_INPUT)
    ids = [s[0] for s in shares]
    if len(set(ids)) != len(ids):
        raise ShareErr(INVALID_INPUT)
    if len(shares) < deg + 1:
        raise ShareErr(INSUFFICIENT_SHARES)
    domain_size(n)
    for i in ids:
        if i >= n:
            raise ShareErr(INVALID_INPUT)
    xs = [domain_element(n, i) for i in ids]
    poly = lagrange_interpolate(xs, [s[1] for s in shares])
    if p_degree(poly) > deg:
        raise ShareErr(DEGREE_MISMATCH)
    # reference indexes result_poly[0]; for the zero polynomial that would panic - we return 0
    return poly, (poly[0] if poly else 0)


# robust_interpolate_fnt: robust_interpolate.rs:206-266.  shares = id-sorted [(id, value)] prefix.
def robust_interpolate_fnt(t: int, n: int, shares, degree: int):
    subset = shares[: degree + 1]
    xs = [domain_element(n, i) for i, _ in subset]
    ys = [y for _, y in subset]
    a_poly = [1]
    for x in xs:
        a_poly = p_mul(a_poly, [(-x) % R_MOD, 1])
    a_der = p_norm([(i * c) % R_MOD for i, c in enumerate(a_poly)][1:])
    interpolated = []
    for i, x_i in enumerate(xs):
        denom = p_eval(a_der, x_i)
        if denom == 0:
            raise ShareErr(POLYNOMIAL_OPERATION_ERROR)
        scalar = (ys[i] * inv(denom)) % R_MOD
        basis, rem = p_divmod(a_poly, [(-x_i) % R_MOD, 1])
        if rem:
            raise ShareErr(POLYNOMIAL_OPERATION_ERROR)
        interpolated = p_add(interpolated, p_scale(basis, scalar))
    valid = sum(1 for i, y in shares if p_eval(interpolated, domain_element(n, i)) == y)
    if valid >= degree + t + 1:
        return interpolated
    raise ShareErr(DECODING_ERROR)


# compute_g0_from_domain: robust_interpolate.rs:540-565
def compute_g0_from_domain(n: int):
    g0 = [1]
    for i in range(n):
        g0 = p_mul(g0, [(-domain_element(n, i)) % R_MOD, 1])
    return g0


# gao_rs_decode: robust_interpolate.rs:456-538
def gao_rs_decode(received, k: int, n: int, erasures):
    if k > n:
        raise ShareErr(INVALID_INPUT)
    s_set = set(erasures)
    s = len(s_set)
    s_poly = [1]
    for i in s_set:
        s_poly = p_mul(s_poly, [(-domain_element(n, i)) % R_MOD, 1])
    known = [(domain_element(n, i), received[i]) for i in range(n) if i not in s_set]
    g1 = lagrange_interpolate([x for x, _ in known], [y for _, y in known])
    g0, _ = p_divmod(compute_g0_from_domain(n), s_poly)
    threshold = (n - s + k) // 2
    r0, r1 = g0, g1
    t0, t1 = [], [1]
    while p_degree(r1) >= threshold:
        q, _ = p_divmod(r0, r1)
        r = p_sub(r0, p_mul(q, r1))
        tt = p_sub(t0, p_mul(q, t1))
        r0, r1 = r1, r
        t0, t1 = t1, tt
    g, v = r1, t1
    quotient, _ = p_divmod(g, v)
    remainder = p_sub(g, p_mul(quotient, v))
    if not remainder and p_degree(quotient) < k:
        return quotient
    raise ShareErr(DECODING_ERROR)


# oec_decode: robust_interpolate.rs:579-628.  Returns (poly, P(0), round r).
def oec_decode(n: int, t: int, shares, degree: int):
    for r in range(1, t + 1):
        required = degree + t + 1 + r
        if len(shares) < required:
            break
        subset = shares[:required]
        received = [0] * n
        erasures = []
        have = dict(subset)
        for i in range(n):
            if i in have:
                received[i] = have[i]
            else:
                erasures.append(i)
        try:
            coeffs = gao_rs_decode(received, degree + 1, n, erasures)
        except ShareErr:
            continue
        poly = p_norm(coeffs)
        matched = sum(1 for i, y in subset if p_eval(poly, domain_element(n, i)) == y)
        if matched >= degree + t + 1:
            return poly, p_eval(poly, 0), r
    raise ShareErr(DECODING_ERROR)


# RobustShare::recover_secret: robust_interpolate.rs:94-157.
# shares = [(id, value, degree)] in arrival order.  Returns dict(coeffs (trimmed), secret, path, flags)
# where path = 0 for the optimistic path, r>0 for OEC round r, and flags[i] says supplied share i
# (arrival order) disagrees with the decoded polynomial (the predicate of :253-257 / :614-617).
def robust_recover_secret(shares, n: int, t: int):
    if n < 3 * t + 1:
        raise ShareErr(INVALID_INPUT)
    if not shares:
        raise ShareErr(INVALID_INPUT)
    degree = shares[0][2]
    if any(s[2] != degree for s in shares):
        raise ShareErr(DEGREE_MISMATCH)
    ids = [s[0] for s in shares]
    if len(set(ids)) != len(ids):
        raise ShareErr(INVALID_INPUT)
    if any(i >= n for i in ids):
        raise ShareErr(INVALID_INPUT)
    if len(shares) < degree + t + 1:
        raise ShareErr(INVALID_INPUT)
    srt = sorted(((s[0], s[1]) for s in shares), key=lambda s: s[0])
    try:
        poly = robust_interpolate_fnt(t, n, srt[: degree + t + 1], degree)
        path = 0
    except ShareErr:
        poly, _, path = oec_decode(n, t, srt, degree)
    flags = [p_eval(poly, domain_element(n, s[0])) != s[1] for s in shares]
    return {"coeffs": poly, "secret": p_eval(poly, 0), "path": path, "flags": flags}


# batch_recover_secret: robust_interpolate.rs:284-443.
# evals_by_sender = [(sender_id, [values per chunk])] in arrival order.
# Returns list per chunk of dict(coeffs (as the reference returns them: fixed d+1 on the optimistic
# path, trimmed on the fallback path), path).
def batch_recover_secret(evals_by_sender, n: int, degree: int, t: int):
    if n < 3 * t + 1:
        raise ShareErr(INVALID_INPUT)
    if not evals_by_sender:
        raise ShareErr(INVALID_INPUT)
    batch_len = len(evals_by_sender[0][1])
    if batch_len == 0:
        raise ShareErr(INVALID_INPUT)
    if any(len(v) != batch_len for _, v in evals_by_sender):
        raise ShareErr(INVALID_INPUT)
    srt = sorted(evals_by_sender, key=lambda e: e[0])
    seen = set()
    for i, _ in srt:
        if i in seen:
            raise ShareErr(INVALID_INPUT)
        seen.add(i)
        if i >= n:
            raise ShareErr(INVALID_INPUT)
    needed = degree + t + 1
    if len(srt) < needed:
        raise ShareErr(INVALID_INPUT)
    m = degree + 1
    xs = [domain_element(n, srt[i][0]) for i in range(m)]
    a_poly = [1]
    for x in xs:
        a_poly = p_mul(a_poly, [(-x) % R_MOD, 1])
    a_der = p_norm([(i * c) % R_MOD for i, c in enumerate(a_poly)][1:])
    basis = []
    for x_i in xs:
        denom = p_eval(a_der, x_i)
        if denom == 0:
            raise ShareErr(POLYNOMIAL_OPERATION_ERROR)
        b, rem = p_divmod(a_poly, [(-x_i) % R_MOD, 1])
        if rem:
            raise ShareErr(POLYNOMIAL_OPERATION_ERROR)
        basis.append(p_scale(b, inv(denom)))
    verify_xs = [domain_element(n, srt[s][0]) for s in range(needed)]
    verify = [[p_eval(basis[i], verify_xs[s]) for i in range(m)] for s in range(needed)]
    out = []
    for c in range(batch_len):
        ok = True
        for s in range(needed):
            acc = sum(verify[s][i] * srt[i][1][c] for i in range(m)) % R_MOD
            if acc != srt[s][1][c]:
                ok = False
                break
        if ok:
            coeffs = [
                sum((basis[i][k] if k < len(basis[i]) else 0) * srt[i][1][c] for i in range(m)) % R_MOD
                for k in range(degree + 1)
            ]
            out.append({"coeffs": coeffs, "path": 0})
        else:
            shares = [(i, v[c], degree) for i, v in srt]
            rec = robust_recover_secret(shares, n, t)  # error propagates like `?` at :437
            out.append({"coeffs": rec["coeffs"], "path": rec["path"]})
    return out


# ---------------------------------------------------------------------------------------------
# U256 boundary helpers: ffi/c_bindings/mod.rs:17-49 (4 x u64 little-endian limbs, canonical)
# ---------------------------------------------------------------------------------------------
def to_limbs(x: int):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


# Deterministic synthetic-input generator shared by tests/bench (SplitMix64 + rejection).
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def fr(self) -> int:
        while True:
            v = self.next() | (self.next() << 64) | (self.next() << 128) | ((self.next() >> 1) << 192)
            if v < R_MOD:
                return v
