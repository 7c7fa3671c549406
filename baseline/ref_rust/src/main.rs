//! Times `Problem::solve()` of the reference crate on this repository's synthetic dense LP families
//! (oracle/synth_lp.hpp, kinds 0..2; the generator below restates it: same splitmix64 hash, same streams), so that the
//! reference's own CPU time can be put next to the GPU engine's on identical inputs:
//!
//!     cargo run --release -- <kind> <m> <n> <seed>
//!
//! The crate exposes no pivot counter (everything below Problem/Solution is private), so this reports seconds to the
//! optimum and the objective; compare with `python -c "import minilp_b200 as mb; ..."` or tests/tools/deep_parity.py,
//! which print pivots and seconds for the engine and for the C++ oracle on the same (kind, m, n, seed).
//!
//! UNVERIFIED: there is no Rust toolchain in the environment this repository was built in.
use minilp::{ComparisonOp, OptimizationDirection, Problem};
use std::time::Instant;

fn mix(mut z: u64) -> u64 {
    z = z.wrapping_add(0x9E3779B97F4A7C15);
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}
fn stream_key(seed: u64, stream: u64) -> u64 {
    mix(seed.wrapping_add(0x632BE59BD9B4E019u64.wrapping_mul(stream + 1)))
}
fn u01(key: u64, idx: u64) -> f64 {
    ((mix(key ^ idx) >> 11) as f64) * (1.0 / 9007199254740992.0) // 2^-53
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() < 5 {
        eprintln!("usage: {} <kind 0|1|2> <m> <n> <seed>", args[0]);
        std::process::exit(2);
    }
    let kind: u32 = args[1].parse().unwrap();
    let m: usize = args[2].parse().unwrap();
    let n: usize = args[3].parse().unwrap();
    let seed: u64 = args[4].parse().unwrap();
    assert!(kind <= 2, "kind 3 (dense_mixed) needs A x0 for its right-hand sides; use kinds 0..2 here");
    let (ka, kc, kb) = (stream_key(seed, 0), stream_key(seed, 1), stream_key(seed, 2));

    let t0 = Instant::now();
    let dir = if kind == 2 { OptimizationDirection::Minimize } else { OptimizationDirection::Maximize };
    let mut p = Problem::new(dir);
    let vars: Vec<_> = (0..n)
        .map(|j| {
            let hi = if kind == 1 { 1.0 } else { f64::INFINITY };
            p.add_var(0.5 + u01(kc, j as u64), (0.0, hi))
        })
        .collect();
    for i in 0..m {
        let row: Vec<_> = (0..n)
            .map(|j| {
                let u = u01(ka, (i * n + j) as u64);
                (vars[j], if kind == 1 { 2.0 * u - 1.0 } else { u })
            })
            .collect();
        let b = match kind {
            1 => 0.25 * (0.5 + u01(kb, i as u64)) * (n as f64).sqrt(),
            _ => (n as f64 / 4.0) * (0.5 + u01(kb, i as u64)),
        };
        let op = if kind == 2 { ComparisonOp::Ge } else { ComparisonOp::Le };
        p.add_constraint(row.as_slice(), op, b);
    }
    let t1 = Instant::now();
    let sol = p.solve();
    let t2 = Instant::now();
    match sol {
        Ok(s) => println!(
            "{{\"kind\": {}, \"m\": {}, \"n\": {}, \"seed\": {}, \"build_s\": {:.3}, \"solve_s\": {:.3}, \"objective\": {:.12e}}}",
            kind, m, n, seed, (t1 - t0).as_secs_f64(), (t2 - t1).as_secs_f64(), s.objective()
        ),
        Err(e) => println!("{{\"error\": \"{}\"}}", e),
    }
}
