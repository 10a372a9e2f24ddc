//! dump_reference_golden.rs -- golden vectors of the share hot path, produced BY THE REFERENCE ITSELF (arkworks 0.5).
//!
//! This image has no Rust toolchain, so nothing in this repo was ever checked against bytes that arkworks produced
//! ("parity unpinned").  This file is the ready-to-run recipe that closes the gap on any machine with cargo:
//!
//!   cp tools/reference_golden/dump_reference_golden.rs  <reference>/mpc/examples/dump_reference_golden.rs
//!   cd <reference>/mpc && cargo run --release --example dump_reference_golden > reference_hotpath.json
//!   cp reference_hotpath.json  <this repo>/tests/golden/reference_hotpath.json
//!   python -m pytest tests/test_reference_golden.py            # CPU: oracle vs reference;  -m gpu: CUDA path vs reference
//!
//! (an `examples/` binary sees the crate's public API and its dependencies: ark-std (whose re-exported rand 0.8 is the generator the
//! reference itself uses: honeybadger/mod.rs:69, robust_interpolate.rs:11), ark-poly, ark-ff and ark-bls12-381 are all already in
//! mpc/Cargo.toml; the JSON is written by hand, so nothing else is needed.  tests/test_reference_golden.py skips while
//! tests/golden/reference_*.json is absent and documents the schema.)
//!
//! What it dumps, and the reference code each case exercises:
//!   * "rng_draws"        StdRng::from_seed(seed) -> Fr::rand x k : pins the device sampler (ChaCha12 + Fp::rand rejection,
//!                        csrc/sampler.cuh) to rand 0.8 / ark-ff 0.5.
//!   * "compute_shares"   RobustShare::compute_shares (robust_interpolate.rs:52-82) with a seeded rng; the coefficients it drew are
//!                        re-drawn from a clone of the rng (DensePolynomial::rand(degree), coefficient 0 <- secret).
//!   * "robust_recover"   RobustShare::recover_secret (robust_interpolate.rs:94-157) on honest shares, on <= t corrupted shares at
//!                        every position pattern the reference's own tests use (robust_interpolate.rs:683-876), on sender subsets,
//!                        on > t errors (DecodingError) -- the shapes of benches/hmpc_mul_micro_bench.rs:37-76 included.
//!   * "batch_recover"    batch_recover_secret (robust_interpolate.rs:284-443), honest and with one corrupted sender
//!                        (benches/hmpc_mul_micro_bench.rs:84-118 shapes).
//!   * "vandermonde"      make_vandermonde + apply_vandermonde (common/share/mod.rs:31-76).
//!   * "nonrobust"        NonRobustShare::compute_shares / recover_secret (common/share/shamir.rs:158-239).
//! Values are 0x-prefixed big-endian hex of the canonical integer (`into_bigint`), the format of tests/golden/hotpath_golden.json.

use ark_bls12_381::Fr;
use ark_ff::{BigInteger, PrimeField, UniformRand};
use ark_poly::{univariate::DensePolynomial, DenseUVPolynomial};
use ark_std::rand::{rngs::StdRng, SeedableRng};
use stoffelcrypto::common::{
    share::{apply_vandermonde, make_vandermonde, shamir::NonRobustShare},
    SecretSharingScheme,
};
use stoffelcrypto::honeybadger::robust_interpolate::{
    robust_interpolate::{batch_recover_secret, RobustShare},
    InterpolateError,
};

/// Minimal JSON writer (serde_json is not a dependency of the reference crate): an object is a list of (key, rendered value).
struct Obj(Vec<(String, String)>);
impl Obj {
    fn new(kind: &str) -> Self {
        Obj(vec![("kind".into(), jstr(kind))])
    }
    fn put(&mut self, k: &str, v: String) -> &mut Self {
        self.0.push((k.into(), v));
        self
    }
    fn render(&self) -> String {
        let inner: Vec<String> = self.0.iter().map(|(k, v)| format!("{}: {}", jstr(k), v)).collect();
        format!("{{{}}}", inner.join(", "))
    }
}
fn jstr(s: &str) -> String {
    format!("\"{}\"", s) // only identifiers and hex strings are written: nothing to escape
}
fn jnum<T: std::fmt::Display>(v: T) -> String {
    format!("{}", v)
}
fn jlist(items: &[String]) -> String {
    format!("[{}]", items.join(", "))
}
fn jnums<T: std::fmt::Display>(v: &[T]) -> String {
    jlist(&v.iter().map(|x| format!("{}", x)).collect::<Vec<_>>())
}

fn hx(v: &Fr) -> String {
    let bytes = v.into_bigint().to_bytes_be();
    let mut s = String::from("0x");
    for b in bytes {
        s.push_str(&format!("{:02x}", b));
    }
    s
}
/// JSON list of hex strings
fn hxs(v: &[Fr]) -> String {
    jlist(&v.iter().map(|x| jstr(&hx(x))).collect::<Vec<_>>())
}
fn seed32(tag: u64) -> [u8; 32] {
    let mut s = [0u8; 32];
    s[..8].copy_from_slice(&tag.to_le_bytes());
    s[8] = 0xB2;
    s
}
/// ShareErrorCode numbering of ffi/c_bindings/share/mod.rs:19-36 / 274-284
fn rc_of(e: &InterpolateError) -> u32 {
    match e {
        InterpolateError::PolynomialOperationError(_) => 7,
        InterpolateError::InvalidInput(_) => 4,
        InterpolateError::DecodingError(_) => 8,
        InterpolateError::NoSuitableDomain(_) => 6,
        InterpolateError::ShareError(_) => 2, // only DegreeMismatch is reachable from recover_secret
    }
}
/// the polynomial compute_shares draws from `rng` for (secret, degree): DensePolynomial::rand(degree), coefficient 0 <- secret
fn drawn_coeffs(secret: Fr, degree: usize, rng: &mut StdRng) -> Vec<Fr> {
    let mut poly = DensePolynomial::<Fr>::rand(degree, rng);
    poly.coeffs[0] = secret;
    let mut c = poly.coeffs.clone();
    c.resize(degree + 1, Fr::from(0u64));
    c
}

fn main() {
    let mut cases: Vec<String> = Vec::new();

    // ---- the sampler
    for tag in [1u64, 2, 3] {
        let seed = seed32(tag);
        let mut rng = StdRng::from_seed(seed);
        let draws: Vec<Fr> = (0..64).map(|_| Fr::rand(&mut rng)).collect();
        let mut o = Obj::new("rng_draws");
        o.put("seed", jnums(&seed)).put("draws", hxs(&draws));
        cases.push(o.render());
    }

    // ---- compute_shares + recover_secret
    let shapes: &[(usize, usize)] = &[(4, 1), (5, 1), (7, 2), (10, 3), (13, 4), (16, 5), (20, 6), (64, 21), (128, 42)];
    for (si, &(n, t)) in shapes.iter().enumerate() {
        for degree in [t, 2 * t] {
            if degree + t + 1 > n {
                continue;
            }
            let seed = seed32(100 + 10 * si as u64 + (degree != t) as u64);
            let mut rng = StdRng::from_seed(seed);
            let secret = Fr::rand(&mut rng);
            let mut rng_clone = rng.clone();
            let shares = RobustShare::<Fr>::compute_shares(secret, n, degree, None, &mut rng).unwrap();
            let coeffs = drawn_coeffs(secret, degree, &mut rng_clone);
            let vals: Vec<Fr> = shares.iter().map(|s| s.share[0]).collect();
            let mut o = Obj::new("compute_shares");
            o.put("n", jnum(n)).put("d", jnum(degree)).put("seed", jnums(&seed)).put("coeffs", hxs(&coeffs)).put("shares", hxs(&vals));
            cases.push(o.render());

            // error patterns: none; the first / last / evenly spread e positions of the supplied list for e = 1..t (the first-e pattern
            // with values += i + 7 is benches/hmpc_mul_micro_bench.rs:41-50); t + 1 errors (must fail at degree = t)
            let mut patterns: Vec<(String, Vec<usize>)> = vec![("honest".into(), vec![])];
            for e in 1..=t {
                patterns.push((format!("first{e}"), (0..e).collect()));
                patterns.push((format!("last{e}"), (n - e..n).collect()));
                patterns.push((format!("spread{e}"), (0..e).map(|i| (i * n) / e).collect()));
            }
            if n > 20 {
                patterns = vec![
                    ("honest".into(), vec![]),
                    ("first1".into(), vec![0]),
                    (format!("first{t}"), (0..t).collect()),
                    (format!("last{t}"), (n - t..n).collect()),
                    (format!("spread{t}"), (0..t).map(|i| (i * n) / t).collect()),
                ];
            }
            patterns.push(("over_t".into(), (0..t + 1).collect()));
            // arrival orders / sender subsets: all n in order, all n reversed, the last degree+t+1 ids, even ids then odd ids
            let mut orders: Vec<(String, Vec<usize>)> = vec![
                ("all".into(), (0..n).collect()),
                ("reversed".into(), (0..n).rev().collect()),
                ("tail".into(), (n - (degree + t + 1)..n).collect()),
            ];
            let mut inter: Vec<usize> = (0..n).step_by(2).collect();
            inter.extend((0..n).skip(1).step_by(2));
            orders.push(("interleaved".into(), inter));
            for (pname, errs) in &patterns {
                for (oname, order) in &orders {
                    let mut sub: Vec<RobustShare<Fr>> = order.iter().map(|&i| shares[i].clone()).collect();
                    for (k, &pos) in errs.iter().enumerate() {
                        if pos < sub.len() {
                            sub[pos].share[0] += Fr::from((k as u64) + 7);
                        }
                    }
                    let ids: Vec<usize> = sub.iter().map(|s| s.id).collect();
                    let values: Vec<Fr> = sub.iter().map(|s| s.share[0]).collect();
                    let mut o = Obj::new("robust_recover");
                    o.put("n", jnum(n)).put("t", jnum(t)).put("d", jnum(degree)).put("ids", jnums(&ids)).put("values", hxs(&values));
                    o.put("pattern", jstr(pname)).put("order", jstr(oname)).put("true_coeffs", hxs(&coeffs));
                    match RobustShare::<Fr>::recover_secret(&sub, n, t) {
                        Ok((c, s)) => {
                            // DensePolynomial trims trailing zeros: `coeffs` may be shorter than d + 1
                            o.put("rc", jnum(0)).put("coeffs", hxs(&c)).put("secret", jstr(&hx(&s)));
                        }
                        Err(e) => {
                            o.put("rc", jnum(rc_of(&e)));
                        }
                    }
                    cases.push(o.render());
                }
            }
        }
    }

    // ---- batch_recover_secret: the micro bench's shape (16 chunks) honest, one corrupted sender, and a reversed sender subset
    for &(n, t) in &[(5usize, 1usize), (10, 3), (20, 6), (64, 21)] {
        let degree = t;
        let chunks = 16usize;
        let mut rng = StdRng::from_seed(seed32(500 + n as u64));
        let mut columns: Vec<Vec<Fr>> = vec![Vec::new(); n]; // columns[id][c]
        let mut truth: Vec<String> = Vec::new();
        for _ in 0..chunks {
            let secret = Fr::rand(&mut rng);
            let mut rng_clone = rng.clone();
            let sh = RobustShare::<Fr>::compute_shares(secret, n, degree, None, &mut rng).unwrap();
            truth.push(hxs(&drawn_coeffs(secret, degree, &mut rng_clone)));
            for s in sh {
                columns[s.id].push(s.share[0]);
            }
        }
        for variant in ["honest", "one_bad_sender", "subset_reversed"] {
            let mut evals: Vec<(usize, Vec<Fr>)> = columns.iter().cloned().enumerate().collect();
            if variant == "one_bad_sender" {
                for v in evals[1].1.iter_mut() {
                    *v += Fr::from(3u64);
                }
            }
            if variant == "subset_reversed" {
                evals.reverse();
                evals.truncate(degree + t + 1);
            }
            let mut o = Obj::new("batch_recover");
            o.put("n", jnum(n)).put("t", jnum(t)).put("d", jnum(degree)).put("variant", jstr(variant));
            o.put("ids", jnums(&evals.iter().map(|e| e.0).collect::<Vec<_>>()));
            o.put("evals", jlist(&evals.iter().map(|e| hxs(&e.1)).collect::<Vec<_>>()));
            o.put("true_coeffs", jlist(&truth));
            match batch_recover_secret(&evals, n, degree, t) {
                Ok(polys) => {
                    o.put("rc", jnum(0)).put("coeffs", jlist(&polys.iter().map(|c| hxs(c)).collect::<Vec<_>>()));
                }
                Err(e) => {
                    o.put("rc", jnum(rc_of(&e)));
                }
            }
            cases.push(o.render());
        }
    }

    // ---- make_vandermonde / apply_vandermonde
    for &(n, t) in &[(4usize, 1usize), (5, 1), (10, 3), (20, 6), (64, 21)] {
        let mut rng = StdRng::from_seed(seed32(700 + n as u64));
        let inputs: Vec<RobustShare<Fr>> = (0..t + 1).map(|_| RobustShare::new(Fr::rand(&mut rng), 0, t)).collect();
        let vm = make_vandermonde::<Fr>(n, t).unwrap();
        let out = apply_vandermonde(&vm, &inputs).unwrap();
        let mut o = Obj::new("vandermonde");
        o.put("n", jnum(n)).put("t", jnum(t)).put("matrix", jlist(&vm.iter().map(|r| hxs(r)).collect::<Vec<_>>()));
        o.put("inputs", hxs(&inputs.iter().map(|s| s.share[0]).collect::<Vec<_>>()));
        o.put("outputs", hxs(&out.iter().map(|s| s.share[0]).collect::<Vec<_>>()));
        cases.push(o.render());
    }

    // ---- NonRobustShare: shares on the same domain points, exact-degree interpolation (common/share/shamir.rs:158-239)
    for &(n, degree) in &[(4usize, 1usize), (7, 2), (10, 6), (16, 10)] {
        let mut rng = StdRng::from_seed(seed32(900 + n as u64));
        let secret = Fr::rand(&mut rng);
        let shares = NonRobustShare::<Fr>::compute_shares(secret, n, degree, None, &mut rng).unwrap();
        let ids: Vec<usize> = shares.iter().map(|s| s.id).collect();
        let vals: Vec<Fr> = shares.iter().map(|s| s.share[0]).collect();
        let mut o = Obj::new("nonrobust");
        o.put("n", jnum(n)).put("d", jnum(degree)).put("secret", jstr(&hx(&secret))).put("ids", jnums(&ids)).put("values", hxs(&vals));
        match NonRobustShare::<Fr>::recover_secret(&shares, n, 0) {
            Ok((c, s)) => {
                o.put("rc", jnum(0)).put("coeffs", hxs(&c)).put("recovered", jstr(&hx(&s)));
            }
            Err(_) => {
                o.put("rc", jnum(1));
            }
        }
        cases.push(o.render());
    }

    println!(
        "{{\"generator\": \"tools/reference_golden/dump_reference_golden.rs run inside the reference crate (arkworks 0.5, ark_std::rand StdRng)\", \"reference\": \"Stoffel-Labs/mpc-protocols, crate stoffelcrypto\", \"cases\": [{}]}}",
        cases.join(",\n")
    );
}
