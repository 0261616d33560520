//! tapstark-gpu -- the reference-side binding of libtapstark_b200.so.
//!
//! UNCOMPILED in the build image (no Rust toolchain); written against Plonky3 rev 72b2fc16 as pinned by
//! the reference's manifests.  Three layers:
//!   * `sys`      : the `extern "C"` declarations of include/tapstark.h used here;
//!   * `GpuDft`   : `p3_dft::TwoAdicSubgroupDft<BabyBear>` -- plugs into an UNMODIFIED
//!                  `TwoAdicFriPcs<Val, Dft, ..>` (fri/src/two_adic_pcs.rs:207, used at :237-240);
//!   * `GpuTwoAdicFriPcs` : `basic::bf_pcs::Pcs` with device-resident LDE, Blake3-Merkle MMCS and the
//!                  device commit phase (what `uni_stark::prove` consumes, uni-stark/src/config.rs:33-62).
//! Status codes != 0 become panics, matching the reference's assert!/expect convention.
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_uint, c_void};
use std::ffi::CStr;

use p3_baby_bear::BabyBear;
use p3_dft::TwoAdicSubgroupDft;
use p3_field::AbstractField;
use p3_matrix::bitrev::{BitReversableMatrix, BitReversedMatrixView};
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;

pub mod sys {
    use super::*;
    #[repr(C)] pub struct ts_ctx { _p: [u8; 0] }
    #[repr(C)] pub struct ts_matrix { _p: [u8; 0] }
    #[repr(C)] pub struct ts_tree { _p: [u8; 0] }
    #[repr(C)] pub struct ts_challenger { _p: [u8; 0] }
    extern "C" {
        pub fn ts_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut ts_ctx) -> c_int;
        pub fn ts_ctx_destroy(ctx: *mut ts_ctx);
        pub fn ts_last_error(ctx: *const ts_ctx) -> *const c_char;
        pub fn ts_coset_lde_batch_host(ctx: *mut ts_ctx, evals: *const u32, rows: usize, width: usize,
                                       added_bits: c_uint, shift_monty: u32, natural_order: c_int,
                                       out: *mut u32) -> c_int;
        pub fn ts_matrix_from_host(ctx: *mut ts_ctx, host: *const u32, rows: usize, width: usize,
                                   out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_matrix_download(ctx: *mut ts_ctx, m: *const ts_matrix, row0: usize, nrows: usize,
                                  host: *mut u32) -> c_int;
        pub fn ts_matrix_free(m: *mut ts_matrix);
        pub fn ts_dft_batch(ctx: *mut ts_ctx, coeffs: *const ts_matrix, out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_pcs_commit_host(ctx: *mut ts_ctx, evals: *const *const u32, rows: *const usize,
                                  widths: *const usize, domain_shifts_monty: *const u32, n: usize,
                                  log_blowup: c_uint, layout: c_int, root: *mut u8,
                                  out: *mut *mut ts_tree) -> c_int;
        pub fn ts_tree_matrix(t: *const ts_tree, i: usize) -> *mut ts_matrix;
        pub fn ts_tree_free(t: *mut ts_tree);
        pub fn ts_mmcs_open_batch(ctx: *mut ts_ctx, t: *const ts_tree, index: usize, rows_out: *mut u32,
                                  path_out: *mut u8) -> c_int;
        pub fn ts_dot_ext_powers(ctx: *mut ts_ctx, m: *const ts_matrix, alpha_monty: *const u32,
                                 out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_fri_commit_phase(ctx: *mut ts_ctx, inputs: *const *mut ts_matrix, n_inputs: usize,
                                   log_blowup: c_uint, chal: *mut ts_challenger, commits: *mut u8,
                                   trees: *mut *mut ts_tree, final_poly: *mut u32, rounds: *mut usize) -> c_int;
        pub fn ts_challenger_new(out: *mut *mut ts_challenger) -> c_int;
        pub fn ts_challenger_observe_digest(c: *mut ts_challenger, digest: *const u8);
        pub fn ts_challenger_sample_ext(c: *mut ts_challenger, out: *mut u32);
        pub fn ts_fri_fold_ext_host(ctx: *mut ts_ctx, input: *const u32, h: usize, beta_monty: *const u32,
                                    out: *mut u32) -> c_int;
        // Pcs::open on the device-resident LDE (fri/src/two_adic_pcs.rs:260-419)
        pub fn ts_inv_denoms(ctx: *mut ts_ctx, log_h: c_uint, z_monty: *const u32, out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_interpolate_low_coset(ctx: *mut ts_ctx, lde: *const ts_matrix, n: usize, z_monty: *const u32,
                                        inv_denoms: *const ts_matrix, ys_out: *mut u32) -> c_int;
        pub fn ts_reduce_opening_acc(ctx: *mut ts_ctx, dot: *const ts_matrix, inv_denoms: *const ts_matrix,
                                     alpha_pow_offset_monty: *const u32, reduced_ys_monty: *const u32,
                                     acc: *mut ts_matrix) -> c_int;
        pub fn ts_challenger_grind(c: *mut ts_challenger, bits: c_uint, ext: c_int, witness: *mut u32) -> c_int;
        // quotient_values of uni_stark::prove (uni-stark/src/prover.rs:122-194) and the device-to-device second commit
        pub fn ts_quotient_values(ctx: *mut ts_ctx, trace_lde: *const ts_matrix, log_n: c_uint,
                                  log_quotient_degree: c_uint, program: *const u32, n_instr: usize,
                                  consts_monty: *const u32, n_consts: usize, public_values_monty: *const u32,
                                  n_public: usize, alpha_monty: *const u32, chunks_out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_pcs_commit(ctx: *mut ts_ctx, evals: *const *mut ts_matrix, domain_shifts_monty: *const u32, n: usize,
                             log_blowup: c_uint, layout: c_int, root: *mut u8, out: *mut *mut ts_tree) -> c_int;
    }
}

/// One CUDA context per prover thread (the reference's callers are single-threaded).
pub struct GpuContext(*mut sys::ts_ctx);
unsafe impl Send for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Self {
        let mut p = core::ptr::null_mut();
        let rc = unsafe { sys::ts_ctx_create(device, core::ptr::null_mut(), &mut p) };
        assert_eq!(rc, 0, "ts_ctx_create failed: no CUDA device (there is no CPU fallback)");
        Self(p)
    }
    fn check(&self, rc: c_int, what: &str) {
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::ts_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("{what}: {msg}");
        }
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) { unsafe { sys::ts_ctx_destroy(self.0) } }
}

thread_local! { static CTX: GpuContext = GpuContext::new(0); }

/// `BabyBear` is `#[repr(transparent)]` over its Montgomery u32 at this Plonky3 revision [MEM]:
/// a `Vec<BabyBear>` crosses the ABI as-is.
fn as_u32(v: &[BabyBear]) -> *const u32 { v.as_ptr() as *const u32 }

/// Drop-in for `Radix2DitParallel` in `TwoAdicFriPcs<Val, Dft, ..>` (uni-stark/tests/fib_air.rs:113,122).
#[derive(Clone, Debug, Default)]
pub struct GpuDft;

impl TwoAdicSubgroupDft<BabyBear> for GpuDft {
    // The library writes the LDE in committed (bit-reversed) order; exposing it as a bit-reversed VIEW makes
    // the PCS's `.bit_reverse_rows().to_row_major_matrix()` (two_adic_pcs.rs:239-240) free, exactly like
    // Radix2DitParallel's output.
    type Evaluations = BitReversedMatrixView<RowMajorMatrix<BabyBear>>;

    fn dft_batch(&self, mat: RowMajorMatrix<BabyBear>) -> Self::Evaluations {
        self.coset_lde_batch(mat, 0, BabyBear::one())
    }

    fn coset_lde_batch(&self, mat: RowMajorMatrix<BabyBear>, added_bits: usize, shift: BabyBear) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let mut out: Vec<BabyBear> = Vec::with_capacity((h << added_bits) * w);
        CTX.with(|c| {
            let shift_monty: u32 = unsafe { core::mem::transmute(shift) };
            let rc = unsafe {
                sys::ts_coset_lde_batch_host(c.0, as_u32(&mat.values), h, w, added_bits as c_uint, shift_monty,
                                             /*natural_order=*/0, out.as_mut_ptr() as *mut u32)
            };
            c.check(rc, "ts_coset_lde_batch_host");
        });
        unsafe { out.set_len((h << added_bits) * w) };
        // `committed` holds row bitrev(i) of the natural-order LDE at row i
        BitReversedMatrixView::new(RowMajorMatrix::new(out, w))
    }
}

/// `Pcs`-level replacement keeping the LDE on the device: see INTEGRATION.md for the full impl sketch
/// (commit = ts_pcs_commit_host; get_evaluations_on_domain = ts_pcs_get_evaluations_on_domain; open = device
/// alpha-reduction + ts_fri_commit_phase + host query phase through ts_mmcs_open_batch).
pub struct GpuTwoAdicFriPcs {
    pub log_blowup: usize,
    pub num_queries: usize,
    pub proof_of_work_bits: usize,
}
