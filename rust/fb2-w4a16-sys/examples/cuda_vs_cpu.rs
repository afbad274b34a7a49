//! The reference's `cuda_vs_cpu` and `cuda_desc_act_vs_cpu_reference` checks (crates/ferrum-quantization/tests/
//! gptq_parity_test.rs:329-405, 580-660) driven from Rust through this crate, std only: the fixture generator
//! (LCG :28-38, make_synthetic :61-93, make_desc_act_g_idx :50-59), the CPU reference (dequant backend/cpu.rs:2283-2315 /
//! native_safetensors.rs:1391-1420, f64-accumulate GEMM cpu.rs:481-492) and the gate (max-abs / max-ref < 0.05) are
//! restated here; the GPU side is `GptqLinear::forward_host`.
//!
//!     FB2_W4A16_DIR=<dir holding libfb2_w4a16.so> cargo run --release --example cuda_vs_cpu
//!
//! NOT COMPILED in the repository it ships in (no Rust toolchain there); the same fixtures run from Python in
//! tests/test_gpu_linear.py::{test_cuda_vs_cpu, test_cuda_desc_act_vs_cpu_reference}.
use fb2_w4a16_sys::GptqLinear;

struct Lcg(u64);
impl Lcg {
    fn u32(&mut self) -> u32 {
        self.0 = self.0.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
        (self.0 >> 33) as u32
    }
    fn f32(&mut self, lo: f32, hi: f32) -> f32 {
        let u = (self.u32() & 0x00FF_FFFF) as f32 / 16_777_216.0;
        lo + u * (hi - lo)
    }
}

/// f32 -> IEEE binary16 bits, round to nearest even (what `half::f16::from_f32` does).
fn f32_to_f16_bits(x: f32) -> u16 {
    let b = x.to_bits();
    let sign = ((b >> 16) & 0x8000) as u16;
    let exp = ((b >> 23) & 0xFF) as i32;
    let man = b & 0x007F_FFFF;
    if exp == 0xFF {
        return sign | 0x7C00 | if man != 0 { 0x0200 } else { 0 };
    }
    let e = exp - 127 + 15;
    if e >= 0x1F {
        return sign | 0x7C00;
    }
    if e <= 0 {
        if e < -10 {
            return sign;
        }
        let m = man | 0x0080_0000;
        let shift = (14 - e) as u32;
        let half = m >> shift;
        let rem = m & ((1u32 << shift) - 1);
        let mid = 1u32 << (shift - 1);
        let round = (rem > mid || (rem == mid && (half & 1) == 1)) as u32;
        return sign | (half + round) as u16;
    }
    let half = ((e as u32) << 10) | (man >> 13);
    let rem = man & 0x1FFF;
    let round = (rem > 0x1000 || (rem == 0x1000 && (half & 1) == 1)) as u32;
    sign | (half + round) as u16
}

fn f16_bits_to_f32(h: u16) -> f32 {
    let sign = ((h as u32) & 0x8000) << 16;
    let exp = ((h >> 10) & 0x1F) as u32;
    let man = (h & 0x03FF) as u32;
    let bits = if exp == 0 {
        if man == 0 {
            sign
        } else {
            // subnormal: man * 2^-24
            return f32::from_bits(sign) + (if sign != 0 { -1.0 } else { 1.0 }) * (man as f32) * 5.960_464_5e-8;
        }
    } else if exp == 0x1F {
        sign | 0x7F80_0000 | (man << 13)
    } else {
        sign | ((exp + 127 - 15) << 23) | (man << 13)
    };
    f32::from_bits(bits)
}

struct Synthetic {
    qweight: Vec<i32>,
    scales: Vec<f32>,
    qzeros: Vec<i32>,
}

fn make_synthetic(k: usize, n: usize, group: usize, seed: u64) -> Synthetic {
    let mut r = Lcg(seed);
    let groups = k / group;
    let qweight = (0..k / 8 * n).map(|_| r.u32() as i32).collect();
    // on-disk scales are f16 (gptq.rs:53): round them so that the GPU's f16 copy is the same number
    let scales = (0..groups * n).map(|_| f16_bits_to_f32(f32_to_f16_bits(r.f32(0.01, 0.1)))).collect();
    let qzeros = (0..groups * (n / 8))
        .map(|_| {
            let mut w = 0u32;
            for bi in 0..8 {
                w |= (r.u32() & 0xF) << (bi * 4);
            }
            w as i32
        })
        .collect();
    Synthetic { qweight, scales, qzeros }
}

/// [n, k] row-major f32, `g = g_idx[k]` or `k / group`.
fn dequant(s: &Synthetic, g_idx: Option<&[i32]>, k: usize, n: usize, group: usize) -> Vec<f32> {
    let mut w = vec![0f32; n * k];
    for kk in 0..k {
        let g = g_idx.map_or(kk / group, |gi| gi[kk] as usize);
        for col in 0..n {
            let q = ((s.qweight[(kk / 8) * n + col] as u32) >> ((kk % 8) * 4)) & 0xF;
            let z = (((s.qzeros[g * (n / 8) + col / 8] as u32) >> ((col % 8) * 4)) & 0xF) + 1;
            w[col * k + kk] = (q as i32 - z as i32) as f32 * s.scales[g * n + col];
        }
    }
    w
}

fn check(name: &str, k: usize, n: usize, m: usize, seed: u64, act_order: bool, x_of: impl Fn(usize) -> f32) {
    let group = 128;
    let syn = make_synthetic(k, n, group, seed);
    let g_idx: Option<Vec<i32>> = act_order.then(|| (0..k).map(|i| (i % (k / group)) as i32).collect());
    let x_bits: Vec<u16> = (0..m * k).map(|i| f32_to_f16_bits(x_of(i))).collect();
    let x: Vec<f32> = x_bits.iter().map(|&b| f16_bits_to_f32(b)).collect();
    let w = dequant(&syn, g_idx.as_deref(), k, n, group);
    let mut reference = vec![0f32; m * n];
    for r in 0..m {
        for col in 0..n {
            let mut acc = 0f64;
            for p in 0..k {
                acc += x[r * k + p] as f64 * w[col * k + p] as f64;
            }
            reference[r * n + col] = acc as f32;
        }
    }
    let mut lin = GptqLinear::load_gptq(&syn.qweight, &syn.scales, &syn.qzeros, g_idx.as_deref(), None, 4, group, k, n, 0)
        .expect("load_gptq");
    assert_eq!((lin.in_features(), lin.out_features(), lin.is_act_order()), (k, n, act_order));
    let mut y_bits = vec![0u16; m * n];
    lin.forward_host(&x_bits, &mut y_bits, m).expect("forward_host");
    let (mut max_abs, mut max_ref) = (0f32, 0f32);
    for (yb, r) in y_bits.iter().zip(reference.iter()) {
        max_abs = max_abs.max((f16_bits_to_f32(*yb) - r).abs());
        max_ref = max_ref.max(r.abs());
    }
    let rel = max_abs / max_ref.max(1e-6);
    println!("{name}: K={k} N={n} m={m} max_abs={max_abs:.3e} max_ref={max_ref:.3e} rel={rel:.3e}");
    assert!(rel < 0.05, "{name}: GPU forward diverges from the CPU reference (rel {rel})");
}

fn main() {
    // gptq_parity_test.rs:329-405
    check("cuda_vs_cpu", 512, 256, 2, 0xC0FFEE, false, |i| (i as f32 * 0.003).cos());
    // gptq_parity_test.rs:580 (act-order through the balanced i % groups pattern)
    check("cuda_desc_act_vs_cpu_reference", 512, 256, 3, 0x5A170A7, true, |i| (i as f32 * 0.0041).sin());
    println!("ok");
}
