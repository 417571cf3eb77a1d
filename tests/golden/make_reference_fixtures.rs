//! make_reference_fixtures.rs -- pins the CPU oracle to the REAL crate.
//!
//! Not compiled in this repository (the build image has no cargo/rustc).  A maintainer with a Rust toolchain runs it
//! against the unmodified reference, fleximeter/aus-rust (`aus` 0.1.7, rustfft ^6.2.0):
//!
//!     python tests/golden/export_reference_inputs.py              # writes tests/golden/reference_inputs.bin
//!     cp tests/golden/make_reference_fixtures.rs  <aus-rust checkout>/examples/
//!     (cd <aus-rust checkout> && cargo run --release --example make_reference_fixtures -- \
//!          <repo>/tests/golden/reference_inputs.bin <repo>/tests/golden/reference_fixtures.bin)
//!     python -m pytest tests/test_reference_fixtures.py -q        # oracle vs the crate's own outputs
//!
//! Inputs are read from a file (not regenerated) so that both sides see bit-identical samples.  Outputs are what the
//! crate's public functions return, flattened:
//!   rstft (spectrum/fft.rs:119), irstft (:204), complex_to_polar_rstft + analyzer per frame (fft_tools.rs:41,
//!   analysis/analyzer.rs:41 -- the loop of analyzer.rs:99-104), mp::stft_analysis (mp.rs:26), generate_window
//!   (window.rs:110), rfftfreq (fft_tools.rs:199), autocorrelation, dc_bias / dbfs_max / energy / zero_crossing_rate, a fft_size
//!   that is not a power of two (1000), overlap_add (fft_tools.rs:69) and the compute_* building blocks (s_a_t.rs:79-290).
//!
//! File format (little endian), both files: b"AUSFIX01", u64 n_arrays, then per array: u64 name_len, name (utf-8),
//! u64 ndim, ndim x u64 dims, prod(dims) x f64.

use std::collections::BTreeMap;
use std::env;
use std::fs::File;
use std::io::{Read, Write};

use aus::analysis::{analyzer, autocorrelation, dbfs_max, dc_bias, energy, zero_crossing_rate, Analysis};
use aus::analysis::{compute_spectral_centroid, compute_spectral_entropy, compute_spectral_flatness, compute_spectral_kurtosis,
                    compute_spectral_roll_off_point, compute_spectral_skewness, compute_spectral_slope,
                    compute_spectral_slope_region, compute_spectral_variance, make_power_spectrum, make_spectrum_pmf};
use aus::spectrum::{complex_to_polar_rstft, irstft, overlap_add, rfftfreq, rstft};
use aus::{generate_window, WindowType};

type Arrays = BTreeMap<String, (Vec<u64>, Vec<f64>)>;

fn read_u64(f: &mut File) -> u64 {
    let mut b = [0u8; 8];
    f.read_exact(&mut b).unwrap();
    u64::from_le_bytes(b)
}

fn read_arrays(path: &str) -> Arrays {
    let mut f = File::open(path).expect("cannot open the input file");
    let mut magic = [0u8; 8];
    f.read_exact(&mut magic).unwrap();
    assert_eq!(&magic, b"AUSFIX01");
    let n = read_u64(&mut f);
    let mut out = Arrays::new();
    for _ in 0..n {
        let name_len = read_u64(&mut f) as usize;
        let mut name = vec![0u8; name_len];
        f.read_exact(&mut name).unwrap();
        let ndim = read_u64(&mut f) as usize;
        let dims: Vec<u64> = (0..ndim).map(|_| read_u64(&mut f)).collect();
        let count: u64 = dims.iter().product();
        let mut raw = vec![0u8; count as usize * 8];
        f.read_exact(&mut raw).unwrap();
        let data: Vec<f64> = raw.chunks_exact(8).map(|c| f64::from_le_bytes(c.try_into().unwrap())).collect();
        out.insert(String::from_utf8(name).unwrap(), (dims, data));
    }
    out
}

fn write_arrays(path: &str, arrays: &Arrays) {
    let mut f = File::create(path).expect("cannot create the output file");
    f.write_all(b"AUSFIX01").unwrap();
    f.write_all(&(arrays.len() as u64).to_le_bytes()).unwrap();
    for (name, (dims, data)) in arrays {
        f.write_all(&(name.len() as u64).to_le_bytes()).unwrap();
        f.write_all(name.as_bytes()).unwrap();
        f.write_all(&(dims.len() as u64).to_le_bytes()).unwrap();
        for d in dims {
            f.write_all(&d.to_le_bytes()).unwrap();
        }
        for v in data {
            f.write_all(&v.to_le_bytes()).unwrap();
        }
    }
}

fn window_of(code: usize) -> WindowType {
    // window.rs:9-15, declaration order = the integer codes of include/aus_b200.h
    match code {
        0 => WindowType::Bartlett,
        1 => WindowType::Blackman,
        2 => WindowType::Hanning,
        3 => WindowType::Hamming,
        _ => WindowType::Rectangular,
    }
}

fn analysis_row(a: &Analysis) -> [f64; 14] {
    // declaration order of analyzer.rs:10-23
    [a.spectral_centroid, a.spectral_variance, a.spectral_skewness, a.spectral_kurtosis, a.spectral_entropy,
     a.spectral_flatness, a.spectral_roll_off_50, a.spectral_roll_off_75, a.spectral_roll_off_90, a.spectral_roll_off_95,
     a.spectral_slope, a.spectral_slope_0_1_khz, a.spectral_slope_1_5_khz, a.spectral_slope_0_5_khz]
}

/// rstft -> complex_to_polar_rstft -> analyzer per frame (analyzer.rs:99-104), general hop / window
fn stft_features(x: &Vec<f64>, fft_size: usize, hop: usize, w: WindowType, sr: u32) -> (Vec<u64>, Vec<f64>) {
    let spec = rstft(x, fft_size, hop, w);
    let (mag, _phase) = complex_to_polar_rstft(&spec);
    let mut out = Vec::with_capacity(mag.len() * 14);
    for row in mag.iter() {
        out.extend_from_slice(&analysis_row(&analyzer(row, fft_size, sr)));
    }
    (vec![mag.len() as u64, 14], out)
}

// `Complex` is rustfft's (spectrum/fft.rs:11); examples of the crate may use the crate's own dependencies
fn flatten_spec(spec: &Vec<Vec<rustfft::num_complex::Complex<f64>>>) -> (Vec<u64>, Vec<f64>) {
    let k = spec[0].len();
    let mut out = Vec::with_capacity(spec.len() * k * 2);
    for row in spec.iter() {
        for c in row.iter() {
            out.push(c.re);
            out.push(c.im);
        }
    }
    (vec![spec.len() as u64, k as u64, 2], out)
}

fn main() {
    let args: Vec<String> = env::args().collect();
    assert!(args.len() == 3, "usage: make_reference_fixtures <reference_inputs.bin> <reference_fixtures.bin>");
    let inp = read_arrays(&args[1]);
    let get = |name: &str| -> Vec<f64> { inp.get(name).unwrap_or_else(|| panic!("missing input {}", name)).1.clone() };
    let sr: u32 = 44100;
    let mut out = Arrays::new();

    // config 1 shape: fft 2048 / hop 512 / Hann
    let x = get("c1_x");
    out.insert("c1_feat".into(), stft_features(&x, 2048, 512, WindowType::Hanning, sr));
    out.insert("c1_spec".into(), flatten_spec(&rstft(&x, 2048, 512, WindowType::Hanning)));

    // mp::stft_analysis: hop = N/2, Hamming (mp.rs:42)
    let mut x = get("mp_x");
    let res = aus::mp::stft_analysis(&mut x, 2048, sr, Some(8));
    let mut flat = Vec::with_capacity(res.len() * 14);
    for a in res.iter() {
        flat.extend_from_slice(&analysis_row(a));
    }
    out.insert("mp_feat".into(), (vec![res.len() as u64, 14], flat));

    // config 2 shape: rstft -> irstft, fft 4096 / hop 1024, Hamming, two channels
    for ch in 0..2 {
        let x = get(&format!("c2_x{}", ch));
        let spec = rstft(&x, 4096, 1024, WindowType::Hamming);
        out.insert(format!("c2_spec{}", ch), flatten_spec(&spec));
        let audio = irstft(&spec, 4096, 1024, WindowType::Hamming).expect("irstft");
        out.insert(format!("c2_audio{}", ch), (vec![audio.len() as u64], audio));
    }

    // config 3 shape: clips, fft 1024 / hop 256, features only
    let x = get("c3_x");
    let lens = get("c3_lens");
    let mut pos = 0usize;
    let mut rows: Vec<f64> = Vec::new();
    let mut n_rows = 0u64;
    for l in lens.iter() {
        let clip: Vec<f64> = x[pos..pos + *l as usize].to_vec();
        let (d, f) = stft_features(&clip, 1024, 256, WindowType::Hanning, sr);
        n_rows += d[0];
        rows.extend(f);
        pos += *l as usize;
    }
    out.insert("c3_feat".into(), (vec![n_rows, 14], rows));

    // config 4 shape: fft 16384 / hop 4096
    let x = get("c4_x");
    out.insert("c4_feat".into(), stft_features(&x, 16384, 4096, WindowType::Hanning, sr));

    // analyzer on fixed magnitude rows
    let (dims, m) = inp.get("an_mag").unwrap().clone();
    let k = dims[1] as usize;
    let mut flat = Vec::new();
    for r in 0..dims[0] as usize {
        let row: Vec<f64> = m[r * k..(r + 1) * k].to_vec();
        flat.extend_from_slice(&analysis_row(&analyzer(&row, 2048, sr)));
    }
    out.insert("an_feat".into(), (vec![dims[0], 14], flat));

    // windows (window.rs:110) and rfftfreq (fft_tools.rs:199)
    for w in 0..5usize {
        for l in [16usize, 17usize] {
            let v = generate_window(window_of(w), l);
            out.insert(format!("win{}_{}", w, l), (vec![l as u64], v));
        }
    }
    let f = rfftfreq(2048, sr);
    out.insert("rfftfreq_2048".into(), (vec![f.len() as u64], f));

    // autocorrelation and the time-domain scalars on the first 2048 samples of c1_x
    let x = get("c1_x");
    let chunk: Vec<f64> = x[..2048].to_vec();
    let ac = autocorrelation(&chunk, 2048).expect("autocorrelation");
    out.insert("acorr_2048".into(), (vec![ac.len() as u64], ac));
    out.insert("time_scalars".into(),
               (vec![4], vec![dc_bias(&x), dbfs_max(&x), energy(&x), zero_crossing_rate(&x, sr)]));

    // a fft_size that is not a power of two (rustfft plans any length): fft 1000 / hop 250 on the first 20 000 samples of c1_x
    let x = get("c1_x");
    let short: Vec<f64> = x[..20000].to_vec();
    out.insert("np2_feat".into(), stft_features(&short, 1000, 250, WindowType::Hanning, sr));
    let spec = rstft(&short, 1000, 250, WindowType::Hamming);
    out.insert("np2_spec".into(), flatten_spec(&spec));
    let audio = irstft(&spec, 1000, 250, WindowType::Hamming).expect("irstft 1000");
    out.insert("np2_audio".into(), (vec![audio.len() as u64], audio));

    // overlap_add (fft_tools.rs:69) on the first 9 x 16 samples of c1_x, hop 4
    let chunks: Vec<Vec<f64>> = (0..9).map(|h| x[h * 16..(h + 1) * 16].to_vec()).collect();
    let ola = overlap_add(&chunks, 16, 4);
    out.insert("ola_9x16_4".into(), (vec![ola.len() as u64], ola));

    // the compute_* building blocks (spectral_analysis_tools.rs:79-290) on the first row of an_mag, chained like analyzer.rs:42-62
    let (dims, m) = inp.get("an_mag").unwrap().clone();
    let k = dims[1] as usize;
    let row: Vec<f64> = m[..k].to_vec();
    let freqs = rfftfreq(2048, sr);
    let p = make_power_spectrum(&row);
    let m_sum: f64 = row.iter().sum();
    let p_sum: f64 = p.iter().sum();
    let pmf = make_spectrum_pmf(&p, p_sum);
    let c = compute_spectral_centroid(&row, &freqs, m_sum);
    let var = compute_spectral_variance(&pmf, &freqs, c);
    let vals = vec![c, var, compute_spectral_skewness(&pmf, &freqs, c, var), compute_spectral_kurtosis(&pmf, &freqs, c, var),
                    compute_spectral_entropy(&pmf), compute_spectral_flatness(&row, m_sum),
                    compute_spectral_roll_off_point(&p, &freqs, p_sum, 0.85), compute_spectral_slope(&p, p_sum),
                    compute_spectral_slope_region(&p, &freqs, 105.0, 852.0, sr)];
    out.insert("compute_row0".into(), (vec![vals.len() as u64], vals));

    write_arrays(&args[2], &out);
    println!("wrote {} arrays to {}", out.len(), args[2]);
}
