//! ref_dump -- runs the reference's own acceptance scenario (examples/fhe-ram.rs:34-177, same seeds) on the
//! FFT64 backend and writes every INPUT and OUTPUT of the evaluation path as raw limbs, so that
//! `tests/test_poulpy_fixtures.py` can replay the inputs through the CPU oracle and the CUDA path and compare the
//! outputs limb for limb.  Because the inputs travel as limbs, the comparison does not depend on the PRNG or on
//! the encryption conventions -- only on the conventions of the evaluation path (oracle/SPEC.md, rows 3-15).
//!
//! NOT COMPILED in the build container (no Rust toolchain, Poulpy absent).  Everything that touches the
//! reference crate is taken from its sources (file:line in the comments).  The three Poulpy accessors used to read
//! limbs are marked `// POULPY ACCESSOR`: if the pinned revision spells them differently, those are the only lines
//! to adapt (each must yield the i64 coefficients of one (column, limb) polynomial).
//!
//! File format (little endian): records `u32 name_len | name | u64 count | count x i64`.  Limb order inside a
//! record: VecZnx index ((limb * cols) + col) * n + coeff (limb 0 most significant), GLWE col 0 = body;
//! GGSW = [row][col_in] GLWE; GGLWE key = [row] GLWE  -- the order of include/fheram.h.
use std::fs::File;
use std::io::{BufWriter, Write};

use poulpy_backend::FFT64Ref as BackendImpl; // the reference's tests use FFT64Ref; swap for FFT64Avx to pin that one
use poulpy_core::layouts::{
    GGLWEInfos, GGSW, GLWE, GLWEAutomorphismKey, GLWEInfos, GLWEPlaintext, GLWESecret, LWEInfos,
    prepared::GLWESecretPrepared,
};
use poulpy_core::{GLWEEncryptSk, GLWETrace};
use poulpy_hal::{
    api::{ScratchOwnedAlloc, ScratchOwnedBorrow},
    layouts::{ScratchOwned, ZnxView},
    source::Source,
};

use fhe_ram::{Address, EvaluationKeys, EvaluationKeysPrepared, Parameters, Ram};
use rand_core::RngCore;

struct Dump(BufWriter<File>);
impl Dump {
    fn rec(&mut self, name: &str, v: &[i64]) {
        self.0.write_all(&(name.len() as u32).to_le_bytes()).unwrap();
        self.0.write_all(name.as_bytes()).unwrap();
        self.0.write_all(&(v.len() as u64).to_le_bytes()).unwrap();
        for x in v {
            self.0.write_all(&x.to_le_bytes()).unwrap();
        }
    }
}

/// limbs of one GLWE in the order ((limb * cols) + col) * n + coeff
fn glwe_limbs(ct: &GLWE<Vec<u8>>) -> Vec<i64> {
    let cols = ct.rank().as_usize() + 1;
    let size = ct.size();
    let mut out = Vec::new();
    for limb in 0..size {
        for col in 0..cols {
            out.extend_from_slice(ct.data().at(col, limb)); // POULPY ACCESSOR: VecZnx (col, limb) -> &[i64]
        }
    }
    out
}
/// [row][col_in] GLWE
fn ggsw_limbs(g: &GGSW<Vec<u8>>) -> Vec<i64> {
    let mut out = Vec::new();
    for row in 0..g.dnum().as_usize() {
        for col in 0..g.rank().as_usize() + 1 {
            out.extend(glwe_limbs(&g.at(row, col).to_owned())); // POULPY ACCESSOR: GGSW (row, col_in) -> GLWE view
        }
    }
    out
}
/// [row] GLWE of a key-switching key of rank 1
fn gglwe_limbs<K: GGLWEInfos>(k: &K, at: impl Fn(usize) -> GLWE<Vec<u8>>) -> Vec<i64> {
    let mut out = Vec::new();
    for row in 0..k.dnum().as_usize() {
        out.extend(glwe_limbs(&at(row)));
    }
    out
}

fn main() {
    // ---- examples/fhe-ram.rs:37-95, verbatim seeds -------------------------------------------------------
    let mut source_xs = Source::new([0u8; 32]);
    let mut source_xa = Source::new([0u8; 32]);
    let mut source_xe = Source::new([0u8; 32]);
    let params: Parameters<BackendImpl> = Parameters::<BackendImpl>::new();
    let module = params.module();
    let mut sk: GLWESecret<Vec<u8>> = GLWESecret::alloc_from_infos(&params.glwe_ct_infos());
    sk.fill_ternary_prob(0.5, &mut source_xs);
    let keys: EvaluationKeys<Vec<u8>> = EvaluationKeys::encrypt_sk(&params, &sk, &mut source_xa, &mut source_xe);
    let mut scratch: ScratchOwned<BackendImpl> = ScratchOwned::alloc(1 << 24);
    let mut sk_prep: GLWESecretPrepared<Vec<u8>, BackendImpl> = GLWESecretPrepared::alloc(module, sk.rank());
    sk_prep.prepare(module, &sk);
    let mut keys_prepared: EvaluationKeysPrepared<Vec<u8>, BackendImpl> = EvaluationKeysPrepared::alloc(&params);
    keys_prepared.prepare(module, &keys, scratch.borrow());
    let mut source = Source::new([5u8; 32]);
    let ws = params.word_size();
    let mut data: Vec<u8> = vec![0u8; params.max_addr() * ws];
    source.fill_bytes(data.as_mut_slice());
    let mut ram: Ram<BackendImpl> = Ram::new();
    ram.encrypt_sk(&data, &sk, &mut source_xa, &mut source_xe);
    let mut addr: Address<Vec<u8>> = Address::alloc_from_params(&params);
    let idx: u32 = source.next_u32() % params.max_addr() as u32;
    addr.encrypt_sk(&params, idx, &sk, &mut source_xa, &mut source_xe, scratch.borrow());

    let mut d = Dump(BufWriter::new(File::create("poulpy_fhe_ram.bin").unwrap()));
    // parameters: src/parameters.rs:11-21 as the snapshot has them
    d.rec(
        "params[log_n,base2k,k_pt,k_ct,k_addr,k_evk_trace,k_evk_ggsw_inv,word_size,max_addr,idx]",
        &[
            12, params.basek() as i64, params.k_glwe_pt().as_usize() as i64, params.k_glwe_ct().as_usize() as i64,
            params.k_ggsw_addr().as_usize() as i64, params.k_evk_trace().as_usize() as i64,
            params.k_evk_ggsw_inv().as_usize() as i64, ws as i64, params.max_addr() as i64, idx as i64,
        ],
    );
    d.rec("decomp_n", &params.decomp_n().iter().map(|x| *x as i64).collect::<Vec<_>>());
    d.rec("data", &data.iter().map(|x| *x as i64).collect::<Vec<_>>());
    // secret (ternary coefficients), only needed for the decrypt-level cross-check
    d.rec("sk", sk.data().at(0, 0)); // POULPY ACCESSOR: ScalarZnx (col, 0) -> &[i64]
    // evaluation keys (src/keys.rs:21-25): trace keys in GLWE::trace_galois_elements order, tsk, atk(-1)
    let gal_els: Vec<i64> = GLWE::trace_galois_elements(module);
    d.rec("gal_els", &gal_els);
    for (i, g) in gal_els.iter().enumerate() {
        let k: &GLWEAutomorphismKey<Vec<u8>> = keys.atk_glwe().get(g).unwrap();
        d.rec(&format!("atk_glwe[{}]", i), &gglwe_limbs(k, |row| k.at(row, 0).to_owned()));
    }
    let inv = keys.atk_ggsw_inv();
    d.rec("atk_ggsw_inv", &gglwe_limbs(inv, |row| inv.at(row, 0).to_owned()));
    let tsk = keys.tsk_ggsw_inv();
    d.rec("tsk_ggsw_inv", &gglwe_limbs(tsk, |row| tsk.at(0, 0).at(row, 0).to_owned())); // rank 1: one (0, 0) GGLWE
    // RAM before anything (SubRam::data, needs subram_getters.patch) and the address (src/address.rs:21-24)
    let dump_ram = |d: &mut Dump, name: &str, ram: &Ram<BackendImpl>| {
        let mut v = Vec::new();
        for s in ram.subrams.iter() {
            for ct in s.data().iter() {
                v.extend(glwe_limbs(ct));
            }
        }
        d.rec(name, &v);
    };
    let dump_tree = |d: &mut Dump, name: &str, ram: &Ram<BackendImpl>| {
        let mut v = Vec::new();
        for s in ram.subrams.iter() {
            if let Some(last) = s.tree().last() {
                v.extend(glwe_limbs(&last[0]));
            }
        }
        d.rec(name, &v);
    };
    let dump_cts = |d: &mut Dump, name: &str, cts: &Vec<GLWE<Vec<u8>>>| {
        let mut v = Vec::new();
        for ct in cts.iter() {
            v.extend(glwe_limbs(ct));
        }
        d.rec(name, &v);
    };
    dump_ram(&mut d, "ram_initial", &ram);
    let mut v = Vec::new();
    for c in addr.coordinates.iter() {
        for g in c.value.iter() {
            v.extend(ggsw_limbs(g));
        }
    }
    d.rec("address", &v);

    // ---- the evaluation path: examples/fhe-ram.rs:97-176 --------------------------------------------------
    let ct = ram.read(&addr, &keys_prepared); // :99
    dump_cts(&mut d, "read", &ct);
    let ct = ram.read_prepare_write(&addr, &keys_prepared); // :119
    dump_cts(&mut d, "read_prepare_write", &ct);
    dump_ram(&mut d, "ram_after_rpw", &ram);
    dump_tree(&mut d, "tree_after_rpw", &ram);
    let mut value: Vec<u8> = vec![0u8; ws];
    source.fill_bytes(value.as_mut_slice()); // :141
    d.rec("value", &value.iter().map(|x| *x as i64).collect::<Vec<_>>());
    let ct_w: Vec<GLWE<Vec<u8>>> = value
        .iter()
        .map(|wi| {
            // examples/fhe-ram.rs:179-210 encrypt_glwe, Source([1; 32]) twice
            let glwe_infos = params.glwe_ct_infos();
            let pt_infos = params.glwe_pt_infos();
            let mut ct_w: GLWE<Vec<u8>> = GLWE::alloc_from_infos(&glwe_infos);
            let mut pt_w: GLWEPlaintext<Vec<u8>> = GLWEPlaintext::alloc_from_infos(&pt_infos);
            pt_w.encode_coeff_i64(*wi as i64, pt_infos.k(), 0);
            let mut scratch: ScratchOwned<BackendImpl> = ScratchOwned::alloc(GLWE::encrypt_sk_tmp_bytes(module, &glwe_infos));
            let mut xa = Source::new([1u8; 32]);
            let mut xe = Source::new([1u8; 32]);
            ct_w.encrypt_sk(module, &pt_w, &sk_prep, &mut xa, &mut xe, scratch.borrow());
            ct_w
        })
        .collect();
    dump_cts(&mut d, "w", &ct_w);
    ram.write(&ct_w, &addr, &keys_prepared); // :152
    dump_ram(&mut d, "ram_after_write", &ram);
    dump_tree(&mut d, "tree_after_write", &ram);
    let ct = ram.read(&addr, &keys_prepared); // :162
    dump_cts(&mut d, "read_back", &ct);
    d.0.flush().unwrap();
    println!("wrote poulpy_fhe_ram.bin (idx = {}): copy it to tests/golden/ of the B200 repo and run pytest tests/test_poulpy_fixtures.py", idx);
}
