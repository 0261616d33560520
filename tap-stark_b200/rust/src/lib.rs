//! tapstark-gpu -- the reference-side binding of libtapstark_b200.so (include/tapstark.h).
//!
//! UNCOMPILED in the build image (no Rust toolchain there); written against Plonky3 rev 72b2fc16 as pinned by the
//! reference's manifests (fri/Cargo.toml:8-31) and against the reference's own traits.  Layers:
//!   * `sys`               the `extern "C"` declarations of include/tapstark.h used here;
//!   * `GpuDft`            `p3_dft::TwoAdicSubgroupDft<BabyBear>` -- plugs into an UNMODIFIED
//!                         `TwoAdicFriPcs<Val, Dft, ..>` (fri/src/two_adic_pcs.rs:207, used at :237-240);
//!   * `GpuBlake3Mmcs`     `basic::mmcs::bf_mmcs::BFMmcs<T>` (basic/src/mmcs/bf_mmcs.rs:17-68) for T = BabyBear and
//!                         T = BabyBear^4: Blake3 row hash + 2-to-1 Blake3 tree on the device;
//!   * `GpuChallenger`     the challenger bounds of `StarkGenericConfig` (uni-stark/src/config.rs:48-51) over the library's
//!                         `BfChallenger` twin, so the transcript lives where `ts_pcs_open` can drive it;
//!   * `GpuTwoAdicFriPcs`  `basic::bf_pcs::Pcs` (basic/src/bf_pcs.rs:19-88): commit = ONE `ts_pcs_commit_host`,
//!                         open = ONE `ts_pcs_open` whose postcard bytes deserialize straight into the proof types below.
//! Status codes != 0 become panics, matching the reference's assert!/expect convention.
//!
//! The Mmcs/Proof types cannot be the reference's `TapTreeMmcs`/`CommitedProof<BO, B>`: a prover built on
//! this crate uses `GpuTwoAdicFriPcs` as `StarkGenericConfig::Pcs` and `GpuChallenger` as its challenger.
#![allow(non_camel_case_types)]

use core::cell::OnceCell;
use core::ffi::{c_char, c_int, c_uint, c_void};
use std::ffi::CStr;

use basic::bf_pcs::{OpenedValues, Pcs, PcsExpr};
use basic::challenger::chan_field::U32;
use basic::challenger::BfGrindingChallenger;
use basic::mmcs::bf_mmcs::BFMmcs;
use p3_baby_bear::BabyBear;
use p3_challenger::{CanObserve, CanSample, CanSampleBits};
use p3_commit::{PolynomialSpace, TwoAdicMultiplicativeCoset};
use p3_dft::TwoAdicSubgroupDft;
use p3_field::extension::BinomialExtensionField;
use p3_field::{AbstractExtensionField, AbstractField, Field, TwoAdicField};
use p3_matrix::bitrev::BitReversedMatrixView;
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;
use p3_util::{log2_strict_usize, reverse_bits_len};
use serde::{Deserialize, Serialize};

pub type Val = BabyBear;
pub type Challenge = BinomialExtensionField<BabyBear, 4>;
/// `[[u8; 4]; 8]`: what the challenger observes (basic/src/challenger/mod.rs:197-223)
pub type Digest = [U32; 8];

pub mod sys {
    use super::*;
    #[repr(C)] pub struct ts_ctx { _p: [u8; 0] }
    #[repr(C)] pub struct ts_matrix { _p: [u8; 0] }
    #[repr(C)] pub struct ts_tree { _p: [u8; 0] }
    #[repr(C)] pub struct ts_challenger { _p: [u8; 0] }
    #[repr(C)] pub struct ts_taptree { _p: [u8; 0] }
    extern "C" {
        pub fn ts_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut ts_ctx) -> c_int;
        pub fn ts_ctx_destroy(ctx: *mut ts_ctx);
        pub fn ts_last_error(ctx: *const ts_ctx) -> *const c_char;
        pub fn ts_host_register(ctx: *mut ts_ctx, host: *const c_void, bytes: usize) -> c_int;
        pub fn ts_host_unregister(ctx: *mut ts_ctx, host: *const c_void) -> c_int;
        // TwoAdicSubgroupDft (host-buffer forms: H2D, transform, D2H inside the call)
        pub fn ts_coset_lde_batch_host(ctx: *mut ts_ctx, evals: *const u32, rows: usize, width: usize,
                                       added_bits: c_uint, shift_monty: u32, natural_order: c_int,
                                       out: *mut u32) -> c_int;
        pub fn ts_dft_batch_host(ctx: *mut ts_ctx, kind: c_int, mat: *const u32, rows: usize, width: usize,
                                 shift_monty: u32, out: *mut u32) -> c_int;
        // BFMmcs
        pub fn ts_matrix_from_host(ctx: *mut ts_ctx, host: *const u32, rows: usize, width: usize,
                                   out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_matrix_download(ctx: *mut ts_ctx, m: *const ts_matrix, row0: usize, nrows: usize,
                                  host: *mut u32) -> c_int;
        pub fn ts_matrix_rows(m: *const ts_matrix) -> usize;
        pub fn ts_matrix_width(m: *const ts_matrix) -> usize;
        pub fn ts_matrix_free(m: *mut ts_matrix);
        pub fn ts_mmcs_commit(ctx: *mut ts_ctx, mats: *const *mut ts_matrix, n: usize, layout: c_int,
                              take_ownership: c_int, root: *mut u8, out: *mut *mut ts_tree) -> c_int;
        pub fn ts_mmcs_open_batch(ctx: *mut ts_ctx, t: *const ts_tree, index: usize, rows_out: *mut u32,
                                  path_out: *mut u8) -> c_int;
        pub fn ts_mmcs_verify_batch(heights: *const usize, widths: *const usize, k: usize, layout: c_int, index: usize,
                                    rows: *const u32, path: *const u8, depth: usize, root: *const u8) -> c_int;
        pub fn ts_tree_num_matrices(t: *const ts_tree) -> usize;
        pub fn ts_tree_matrix(t: *const ts_tree, i: usize) -> *mut ts_matrix;
        pub fn ts_tree_depth(t: *const ts_tree) -> usize;
        pub fn ts_tree_free(t: *mut ts_tree);
        // Pcs
        pub fn ts_pcs_commit_host(ctx: *mut ts_ctx, evals: *const *const u32, rows: *const usize,
                                  widths: *const usize, domain_shifts_monty: *const u32, n: usize,
                                  log_blowup: c_uint, layout: c_int, root: *mut u8,
                                  out: *mut *mut ts_tree) -> c_int;
        pub fn ts_pcs_get_evaluations_on_domain(ctx: *mut ts_ctx, t: *const ts_tree, idx: usize, domain_size: usize,
                                                out_host: *mut u32) -> c_int;
        pub fn ts_pcs_open(ctx: *mut ts_ctx, rounds: *const *const ts_tree, n_rounds: usize, n_points: *const usize,
                           points_monty: *const u32, log_blowup: c_uint, num_queries: c_uint,
                           proof_of_work_bits: c_uint, chal: *mut ts_challenger, out_bytes: *mut *mut u8,
                           out_len: *mut usize) -> c_int;
        pub fn ts_bytes_free(bytes: *mut u8);
        // fold_even_odd (fri/src/fold_even_odd.rs:20-52)
        pub fn ts_fri_fold_ext_host(ctx: *mut ts_ctx, input: *const u32, h: usize, beta_monty: *const u32,
                                    out: *mut u32) -> c_int;
        // BfChallenger
        pub fn ts_challenger_new(out: *mut *mut ts_challenger) -> c_int;
        pub fn ts_challenger_clone(c: *const ts_challenger, out: *mut *mut ts_challenger) -> c_int;
        pub fn ts_challenger_free(c: *mut ts_challenger);
        pub fn ts_challenger_observe(c: *mut ts_challenger, word: *const u8);
        pub fn ts_challenger_observe_digest(c: *mut ts_challenger, digest: *const u8);
        pub fn ts_challenger_sample_ext(c: *mut ts_challenger, out: *mut u32);
        pub fn ts_challenger_sample_bits(c: *mut ts_challenger, bits: c_uint, ext: c_int) -> usize;
        pub fn ts_challenger_check_witness(c: *mut ts_challenger, bits: c_uint, witness: u32, ext: c_int) -> c_int;
        pub fn ts_challenger_grind(c: *mut ts_challenger, bits: c_uint, ext: c_int, witness: *mut u32) -> c_int;
        // TapTree commitment (basic/src/tcs): leaf scripts from a template, sorted-pair TapBranch tree
        pub fn ts_matrix_from_host(ctx: *mut ts_ctx, host: *const u32, rows: usize, width: usize, out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_matrix_free(m: *mut ts_matrix);
        pub fn ts_padded_leaf_rows(ctx: *mut ts_ctx, mats: *const *const ts_matrix, n_mats: usize, out: *mut *mut ts_matrix) -> c_int;
        pub fn ts_taptree_commit(ctx: *mut ts_ctx, leaf_rows: *const ts_matrix, segs: *const u8, seg_offsets: *const usize,
                                 push_word: *const u32, n_push: usize, root: *mut u8, out: *mut *mut ts_taptree) -> c_int;
        pub fn ts_taptree_leaf_indices(ctx: *mut ts_ctx, t: *const ts_taptree, out_host: *mut u32) -> c_int;
        pub fn ts_taptree_open(ctx: *mut ts_ctx, t: *const ts_taptree, index: usize, path_out: *mut u8, position_out: *mut u32) -> c_int;
        pub fn ts_taptree_free(t: *mut ts_taptree);
    }
}

pub const LAYOUT_P3_INJECT: c_int = 0;

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
    /// Page-locks a long-lived host buffer (a trace `Vec<BabyBear>`) so the `*_host` entry points copy it at the
    /// PCIe rate; pageable sources work too, staged through the library's bounce buffers (bench.py reports both).
    pub fn pin<T>(&self, v: &[T]) {
        let rc = unsafe { sys::ts_host_register(self.0, v.as_ptr() as *const c_void, core::mem::size_of_val(v)) };
        self.check(rc, "ts_host_register");
    }
    pub fn unpin<T>(&self, v: &[T]) {
        let rc = unsafe { sys::ts_host_unregister(self.0, v.as_ptr() as *const c_void) };
        self.check(rc, "ts_host_unregister");
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) { unsafe { sys::ts_ctx_destroy(self.0) } }
}

thread_local! { static CTX: GpuContext = GpuContext::new(0); }

/// `BabyBear` is `#[repr(transparent)]` over its Montgomery u32 at this Plonky3 revision [MEM]: a `Vec<BabyBear>`
/// crosses the ABI as-is, and `BinomialExtensionField<BabyBear, 4>` is `[BabyBear; 4]`.
fn as_u32(v: &[BabyBear]) -> *const u32 { v.as_ptr() as *const u32 }
fn monty(x: BabyBear) -> u32 { unsafe { core::mem::transmute(x) } }
fn digest_bytes(d: &Digest) -> [u8; 32] {
    let mut b = [0u8; 32];
    for (i, w) in d.iter().enumerate() { b[4 * i..4 * i + 4].copy_from_slice(w); }
    b
}
fn digest_from(b: &[u8; 32]) -> Digest { core::array::from_fn(|i| [b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]]) }

// ------------------------------------------------------------------------------------------------ TwoAdicSubgroupDft
/// Drop-in for `Radix2DitParallel` in `TwoAdicFriPcs<Val, Dft, ..>` (uni-stark/tests/fib_air.rs:113,122).
#[derive(Clone, Debug, Default)]
pub struct GpuDft;

const DFT: c_int = 0;
const IDFT: c_int = 1;
const COSET_DFT: c_int = 2;

impl GpuDft {
    fn host_transform(kind: c_int, mat: RowMajorMatrix<BabyBear>, shift: BabyBear) -> RowMajorMatrix<BabyBear> {
        let (h, w) = (mat.height(), mat.width());
        let mut out: Vec<BabyBear> = Vec::with_capacity(h * w);
        CTX.with(|c| {
            let rc = unsafe {
                sys::ts_dft_batch_host(c.0, kind, as_u32(&mat.values), h, w, monty(shift), out.as_mut_ptr() as *mut u32)
            };
            c.check(rc, "ts_dft_batch_host");
        });
        unsafe { out.set_len(h * w) };
        RowMajorMatrix::new(out, w)
    }
}

impl TwoAdicSubgroupDft<BabyBear> for GpuDft {
    // The library writes the LDE in committed (bit-reversed) order; exposing it as a bit-reversed VIEW makes the PCS's
    // `.bit_reverse_rows().to_row_major_matrix()` (two_adic_pcs.rs:239-240) free, exactly like Radix2DitParallel's
    // output.  The plain transforms return natural order wrapped the same way: `BitReversedMatrixView::new(m)` shows
    // row bitrev(i) of m at i, so m is stored pre-reversed (`reversed`) and the view is the natural-order result.
    type Evaluations = BitReversedMatrixView<RowMajorMatrix<BabyBear>>;

    /// out[i] = sum_k coeffs[k] w^(ik): the DFT of the COEFFICIENTS (fri/src/fold_even_odd.rs:75-81 relies on it).
    fn dft_batch(&self, mat: RowMajorMatrix<BabyBear>) -> Self::Evaluations {
        BitReversedMatrixView::new(reversed(Self::host_transform(DFT, mat, BabyBear::one())))
    }
    fn idft_batch(&self, mat: RowMajorMatrix<BabyBear>) -> RowMajorMatrix<BabyBear> {
        Self::host_transform(IDFT, mat, BabyBear::one())
    }
    fn coset_dft_batch(&self, mat: RowMajorMatrix<BabyBear>, shift: BabyBear) -> Self::Evaluations {
        BitReversedMatrixView::new(reversed(Self::host_transform(COSET_DFT, mat, shift)))
    }
    fn coset_lde_batch(&self, mat: RowMajorMatrix<BabyBear>, added_bits: usize, shift: BabyBear) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        let mut out: Vec<BabyBear> = Vec::with_capacity((h << added_bits) * w);
        CTX.with(|c| {
            let rc = unsafe {
                sys::ts_coset_lde_batch_host(c.0, as_u32(&mat.values), h, w, added_bits as c_uint, monty(shift),
                                             /*natural_order=*/0, out.as_mut_ptr() as *mut u32)
            };
            c.check(rc, "ts_coset_lde_batch_host");
        });
        unsafe { out.set_len((h << added_bits) * w) };
        // `out` holds row bitrev(i) of the natural-order LDE at row i: the view is the natural-order LDE
        BitReversedMatrixView::new(RowMajorMatrix::new(out, w))
    }
}
/// rows permuted by bit reversal (host side; only the plain trait calls, which are off the hot path, use it)
fn reversed(m: RowMajorMatrix<BabyBear>) -> RowMajorMatrix<BabyBear> {
    let (h, w) = (m.height(), m.width());
    let lh = log2_strict_usize(h);
    let mut v = vec![BabyBear::zero(); h * w];
    for r in 0..h {
        let s = reverse_bits_len(r, lh);
        v[s * w..(s + 1) * w].copy_from_slice(&m.values[r * w..(r + 1) * w]);
    }
    RowMajorMatrix::new(v, w)
}

// ------------------------------------------------------------------------------------------------ BFMmcs
/// Opening proof of the Blake3 Merkle MMCS.  The reference's `verify_batch` takes no index or dimensions (its Taproot
/// proof carries them, basic/src/tcs/mod.rs:100-106), so this one carries them too.
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct MerkleProof {
    pub index: usize,
    pub heights: Vec<usize>,
    pub siblings: Vec<[u8; 32]>,
}

/// `BFMmcs::ProverData`: the device-resident tree and matrices; host copies are fetched on first `get_matrices`.
pub struct GpuProverData<T> {
    tree: *mut sys::ts_tree,
    heights: Vec<usize>,
    widths: Vec<usize>, // in T elements
    host: OnceCell<Vec<RowMajorMatrix<T>>>,
}
unsafe impl<T: Send> Send for GpuProverData<T> {}
impl<T> Drop for GpuProverData<T> {
    fn drop(&mut self) { unsafe { sys::ts_tree_free(self.tree) } }
}
impl<T> GpuProverData<T> {
    fn from_tree(tree: *mut sys::ts_tree, elem_words: usize) -> Self {
        let k = unsafe { sys::ts_tree_num_matrices(tree) };
        let (mut heights, mut widths) = (Vec::new(), Vec::new());
        for i in 0..k {
            let m = unsafe { sys::ts_tree_matrix(tree, i) };
            heights.push(unsafe { sys::ts_matrix_rows(m) });
            widths.push(unsafe { sys::ts_matrix_width(m) } / elem_words);
        }
        Self { tree, heights, widths, host: OnceCell::new() }
    }
    pub fn raw(&self) -> *const sys::ts_tree { self.tree }
}

/// Field elements the MMCS can commit to: BabyBear (1 word) and BabyBear^4 (4 words, low coefficient first).
pub trait DeviceElem: Copy + Send + Sync + 'static {
    const WORDS: usize;
}
impl DeviceElem for BabyBear { const WORDS: usize = 1; }
impl DeviceElem for Challenge { const WORDS: usize = 4; }

#[derive(Clone, Debug, Default)]
pub struct GpuBlake3Mmcs;

#[derive(Debug)]
pub struct RootMismatch;

impl<T: DeviceElem> BFMmcs<T> for GpuBlake3Mmcs {
    type ProverData = GpuProverData<T>;
    type Commitment = Digest;
    type Proof = MerkleProof;
    type Error = RootMismatch;

    fn commit(&self, inputs: Vec<RowMajorMatrix<T>>) -> (Self::Commitment, Self::ProverData) {
        CTX.with(|c| {
            let mut mats: Vec<*mut sys::ts_matrix> = Vec::with_capacity(inputs.len());
            for m in &inputs {
                let mut h = core::ptr::null_mut();
                let rc = unsafe {
                    sys::ts_matrix_from_host(c.0, m.values.as_ptr() as *const u32, m.height(), m.width() * T::WORDS, &mut h)
                };
                c.check(rc, "ts_matrix_from_host");
                mats.push(h);
            }
            let (mut root, mut tree) = ([0u8; 32], core::ptr::null_mut());
            let rc = unsafe {
                sys::ts_mmcs_commit(c.0, mats.as_ptr(), mats.len(), LAYOUT_P3_INJECT, /*take_ownership=*/1, root.as_mut_ptr(), &mut tree)
            };
            c.check(rc, "ts_mmcs_commit");
            let pd = GpuProverData::from_tree(tree, T::WORDS);
            let _ = pd.host.set(inputs); // the caller's matrices ARE the committed ones: no download needed later
            (digest_from(&root), pd)
        })
    }

    /// `query_times_index` selects one of the reference's `num_queries` Taptrees; a Merkle commitment is one tree.
    fn open_batch(&self, _query_times_index: usize, query_index: usize, pd: &Self::ProverData) -> (Vec<Vec<T>>, Self::Proof) {
        CTX.with(|c| {
            let total: usize = pd.widths.iter().sum::<usize>() * T::WORDS;
            let depth = unsafe { sys::ts_tree_depth(pd.tree) };
            let mut rows = vec![0u32; total];
            let mut path = vec![0u8; 32 * depth.max(1)];
            let rc = unsafe { sys::ts_mmcs_open_batch(c.0, pd.tree, query_index, rows.as_mut_ptr(), path.as_mut_ptr()) };
            c.check(rc, "ts_mmcs_open_batch");
            let mut out = Vec::with_capacity(pd.widths.len());
            let mut o = 0;
            for &w in &pd.widths {
                let words = &rows[o..o + w * T::WORDS];
                o += w * T::WORDS;
                // SAFETY: T is BabyBear or [BabyBear; 4], both transparent over Montgomery u32 words
                out.push(unsafe { core::slice::from_raw_parts(words.as_ptr() as *const T, w) }.to_vec());
            }
            let siblings = (0..depth).map(|l| core::array::from_fn(|i| path[32 * l + i])).collect();
            (out, MerkleProof { index: query_index, heights: pd.heights.clone(), siblings })
        })
    }

    fn verify_batch(&self, _query_times_index: usize, opened_values: &Vec<Vec<T>>, proof: &Self::Proof, root: &Self::Commitment)
        -> Result<(), Self::Error> {
        let widths: Vec<usize> = opened_values.iter().map(|r| r.len() * T::WORDS).collect();
        let flat: Vec<u32> = opened_values.iter()
            .flat_map(|r| unsafe { core::slice::from_raw_parts(r.as_ptr() as *const u32, r.len() * T::WORDS) }.iter().copied())
            .collect();
        let path: Vec<u8> = proof.siblings.iter().flatten().copied().collect();
        let rc = unsafe {
            sys::ts_mmcs_verify_batch(proof.heights.as_ptr(), widths.as_ptr(), widths.len(), LAYOUT_P3_INJECT, proof.index,
                                      flat.as_ptr(), if path.is_empty() { [0u8; 32].as_ptr() } else { path.as_ptr() },
                                      proof.siblings.len(), digest_bytes(root).as_ptr())
        };
        if rc == 0 { Ok(()) } else { Err(RootMismatch) }
    }

    fn get_matrices<'a>(&self, pd: &'a Self::ProverData) -> Vec<&'a RowMajorMatrix<T>> {
        pd.host.get_or_init(|| CTX.with(|c| {
            (0..pd.heights.len()).map(|i| {
                let m = unsafe { sys::ts_tree_matrix(pd.tree, i) };
                let (h, w) = (pd.heights[i], pd.widths[i]);
                let mut v: Vec<T> = Vec::with_capacity(h * w);
                let rc = unsafe { sys::ts_matrix_download(c.0, m, 0, h, v.as_mut_ptr() as *mut u32) };
                c.check(rc, "ts_matrix_download");
                unsafe { v.set_len(h * w) };
                RowMajorMatrix::new(v, w)
            }).collect()
        })).iter().collect()
    }
}

// ------------------------------------------------------------------------------------------------ challenger
/// `BfChallenger<Challenge, U32, Blake3Permutation, 16>` (basic/src/challenger/mod.rs) kept inside the library
/// (csrc/host_side.h restates it; golden 1103171332 of script_expr/src/challenger_expr.rs:278-296 reproduced through
/// this ABI in tests/).  The permutation/sample *records* the reference keeps for script generation are not kept.
pub struct GpuChallenger(*mut sys::ts_challenger);
unsafe impl Send for GpuChallenger {}
unsafe impl Sync for GpuChallenger {}
impl GpuChallenger {
    pub fn new() -> Self {
        let mut p = core::ptr::null_mut();
        assert_eq!(unsafe { sys::ts_challenger_new(&mut p) }, 0);
        Self(p)
    }
}
impl Default for GpuChallenger { fn default() -> Self { Self::new() } }
impl Clone for GpuChallenger {
    fn clone(&self) -> Self {
        let mut p = core::ptr::null_mut();
        assert_eq!(unsafe { sys::ts_challenger_clone(self.0, &mut p) }, 0);
        Self(p)
    }
}
impl Drop for GpuChallenger { fn drop(&mut self) { unsafe { sys::ts_challenger_free(self.0) } } }
impl CanObserve<U32> for GpuChallenger {
    fn observe(&mut self, value: U32) { unsafe { sys::ts_challenger_observe(self.0, value.as_ptr()) } }
}
impl CanObserve<Digest> for GpuChallenger {
    fn observe(&mut self, value: Digest) { unsafe { sys::ts_challenger_observe_digest(self.0, digest_bytes(&value).as_ptr()) } }
}
impl CanSample<Challenge> for GpuChallenger {
    fn sample(&mut self) -> Challenge {
        let mut c = [0u32; 4]; // canonical transcript values (chan_field.rs:12-18: u32 LE mod p)
        unsafe { sys::ts_challenger_sample_ext(self.0, c.as_mut_ptr()) };
        Challenge::from_base_slice(&c.map(BabyBear::from_canonical_u32))
    }
}
impl CanSampleBits<usize> for GpuChallenger {
    fn sample_bits(&mut self, bits: usize) -> usize { unsafe { sys::ts_challenger_sample_bits(self.0, bits as c_uint, 1) } }
}
impl BfGrindingChallenger for GpuChallenger {
    type Witness = U32;
    /// Deterministic: the smallest valid witness (the reference's rayon `find_any` returns any of ~32).
    fn grind(&mut self, bits: usize) -> U32 {
        let mut w = 0u32;
        let rc = unsafe { sys::ts_challenger_grind(self.0, bits as c_uint, 1, &mut w) };
        assert_eq!(rc, 0, "failed to find witness"); // basic/src/challenger/mod.rs:101
        w.to_le_bytes()
    }
    fn check_witness(&mut self, bits: usize, witness: U32) -> bool {
        unsafe { sys::ts_challenger_check_witness(self.0, bits as c_uint, u32::from_le_bytes(witness), 1) != 0 }
    }
}

// ------------------------------------------------------------------------------------------------ proof types
// Field order and nesting = the postcard layout ts_pcs_open emits (include/tapstark.h), which is the serde derive
// order of fri/src/proof.rs:13-33 with the Merkle sibling path in place of `CommitedProof<BO, B>`.
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct BatchOpening {
    pub opened_values: Vec<Vec<Val>>,
    pub opening_proof: Vec<[u8; 32]>,
}
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct BfQueryProof {
    pub input_proof: Vec<BatchOpening>,
    pub commit_phase_openings: Vec<(Vec<Vec<Challenge>>, Vec<[u8; 32]>)>,
}
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct FriProof {
    pub commit_phase_commits: Vec<[u8; 32]>,
    pub query_proofs: Vec<BfQueryProof>,
    pub final_poly: Challenge,
    pub pow_witness: u32,
}

#[derive(Debug)]
pub enum VerifyError {
    InvalidProofShape,
    InvalidPowWitness,
    InputMmcs(usize),
    CommitPhaseMmcs(usize),
    FinalPolyMismatch,
}

// ------------------------------------------------------------------------------------------------ Pcs
#[derive(Clone, Debug)]
pub struct GpuTwoAdicFriPcs {
    pub log_blowup: usize,
    pub num_queries: usize,
    pub proof_of_work_bits: usize,
    pub mmcs: GpuBlake3Mmcs,
}

impl Pcs<Challenge, GpuChallenger> for GpuTwoAdicFriPcs {
    type Domain = TwoAdicMultiplicativeCoset<Val>;
    type Commitment = Digest;
    type ProverData = GpuProverData<Val>;
    type Proof = FriProof;
    type Error = VerifyError;

    fn natural_domain_for_degree(&self, degree: usize) -> Self::Domain {
        TwoAdicMultiplicativeCoset { log_n: log2_strict_usize(degree), shift: Val::one() } // two_adic_pcs.rs:219-225
    }

    /// two_adic_pcs.rs:227-245 in one call: per matrix shift = g / domain.shift, coset LDE, committed order, one MMCS
    /// commit over all of them; the LDEs never leave the device.
    fn commit(&self, evaluations: Vec<(Self::Domain, RowMajorMatrix<Val>)>) -> (Self::Commitment, Self::ProverData) {
        let (mut ptrs, mut rows, mut widths, mut shifts) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (domain, evals) in &evaluations {
            assert_eq!(domain.size(), evals.height()); // :234
            ptrs.push(as_u32(&evals.values));
            rows.push(evals.height());
            widths.push(evals.width());
            shifts.push(monty(domain.shift));
        }
        CTX.with(|c| {
            let (mut root, mut tree) = ([0u8; 32], core::ptr::null_mut());
            let rc = unsafe {
                sys::ts_pcs_commit_host(c.0, ptrs.as_ptr(), rows.as_ptr(), widths.as_ptr(), shifts.as_ptr(), ptrs.len(),
                                        self.log_blowup as c_uint, LAYOUT_P3_INJECT, root.as_mut_ptr(), &mut tree)
            };
            c.check(rc, "ts_pcs_commit_host");
            (digest_from(&root), GpuProverData::from_tree(tree, 1))
        })
    }

    /// two_adic_pcs.rs:247-258: the first `domain.size()` committed rows, re-bit-reversed (done on the device).
    fn get_evaluations_on_domain<'a>(&self, pd: &'a Self::ProverData, idx: usize, domain: Self::Domain) -> impl Matrix<Val> + 'a {
        assert_eq!(domain.shift, Val::generator());
        assert!(pd.heights[idx] >= domain.size());
        let w = pd.widths[idx];
        let mut v: Vec<Val> = Vec::with_capacity(domain.size() * w);
        CTX.with(|c| {
            let rc = unsafe { sys::ts_pcs_get_evaluations_on_domain(c.0, pd.tree, idx, domain.size(), v.as_mut_ptr() as *mut u32) };
            c.check(rc, "ts_pcs_get_evaluations_on_domain");
        });
        unsafe { v.set_len(domain.size() * w) };
        RowMajorMatrix::new(v, w)
    }

    /// two_adic_pcs.rs:260-419 + fri/src/prover.rs:19-90 in one call; the bytes deserialize into the proof types.
    fn open(&self, rounds: Vec<(&Self::ProverData, Vec<Vec<Challenge>>)>, challenger: &mut GpuChallenger)
        -> (OpenedValues<Challenge>, Self::Proof) {
        let trees: Vec<*const sys::ts_tree> = rounds.iter().map(|(pd, _)| pd.raw()).collect();
        let (mut counts, mut pts): (Vec<usize>, Vec<u32>) = (Vec::new(), Vec::new());
        for (pd, points) in &rounds {
            assert_eq!(points.len(), pd.heights.len(), "one point list per committed matrix");
            for per_mat in points {
                counts.push(per_mat.len());
                for z in per_mat { pts.extend(z.as_base_slice().iter().map(|&x| monty(x))); }
            }
        }
        let bytes = CTX.with(|c| {
            let (mut buf, mut n) = (core::ptr::null_mut(), 0usize);
            let rc = unsafe {
                sys::ts_pcs_open(c.0, trees.as_ptr(), trees.len(), counts.as_ptr(), pts.as_ptr(), self.log_blowup as c_uint,
                                 self.num_queries as c_uint, self.proof_of_work_bits as c_uint, challenger.0, &mut buf, &mut n)
            };
            c.check(rc, "ts_pcs_open");
            let v = unsafe { core::slice::from_raw_parts(buf, n) }.to_vec();
            unsafe { sys::ts_bytes_free(buf) };
            v
        });
        postcard::from_bytes::<(OpenedValues<Challenge>, FriProof)>(&bytes).expect("ts_pcs_open: malformed proof bytes")
    }

    /// two_adic_pcs.rs:421-530 + fri/src/verifier.rs:20-165 over the Merkle MMCS (host arithmetic, as in the reference).
    fn verify(&self, rounds: Vec<(Self::Commitment, Vec<(Self::Domain, Vec<(Challenge, Vec<Challenge>)>)>)>, proof: &Self::Proof,
              challenger: &mut GpuChallenger) -> Result<(), Self::Error> {
        let alpha: Challenge = challenger.sample(); // :436
        let log_global_max_height = proof.commit_phase_commits.len() + self.log_blowup; // :438
        let betas: Vec<Challenge> = proof.commit_phase_commits.iter().map(|c| {
            challenger.observe(digest_from(c)); // verifier.rs:36-42
            challenger.sample()
        }).collect();
        if proof.query_proofs.len() != self.num_queries { return Err(VerifyError::InvalidProofShape); }
        if !challenger.check_witness(self.proof_of_work_bits, proof.pow_witness.to_le_bytes()) { // verifier.rs:49-51
            return Err(VerifyError::InvalidPowWitness);
        }
        let g = Val::generator();
        for (qi, qp) in proof.query_proofs.iter().enumerate() {
            let index = challenger.sample_bits(log_global_max_height);
            // reduced openings per log height (:448-500)
            let mut ro = vec![Challenge::zero(); 32];
            let mut alpha_pow = vec![Challenge::one(); 32];
            if qp.input_proof.len() != rounds.len() { return Err(VerifyError::InvalidProofShape); }
            for (bo, (commit, mats)) in qp.input_proof.iter().zip(&rounds) {
                let heights: Vec<usize> = mats.iter().map(|(d, _)| d.size() << self.log_blowup).collect();
                let log_max = log2_strict_usize(*heights.iter().max().ok_or(VerifyError::InvalidProofShape)?);
                let reduced_index = index >> (log_global_max_height - log_max);
                let mp = MerkleProof { index: reduced_index, heights, siblings: bo.opening_proof.clone() };
                BFMmcs::<Val>::verify_batch(&self.mmcs, 0, &bo.opened_values, &mp, commit).map_err(|_| VerifyError::InputMmcs(qi))?;
                for (row, (dom, pts)) in bo.opened_values.iter().zip(mats) {
                    let log_h = log2_strict_usize(dom.size()) + self.log_blowup;
                    let rev = reverse_bits_len(index >> (log_global_max_height - log_h), log_h);
                    let x = g * Val::two_adic_generator(log_h).exp_u64(rev as u64); // :476-478
                    for (z, ps_at_z) in pts {
                        for (&p_at_x, &p_at_z) in row.iter().zip(ps_at_z) { // :480-486
                            ro[log_h] += alpha_pow[log_h] * (-p_at_z + p_at_x) * (-*z + x).inverse();
                            alpha_pow[log_h] *= alpha;
                        }
                    }
                }
            }
            // fri/src/verifier.rs:100-165 verify_query
            if qp.commit_phase_openings.len() != betas.len() { return Err(VerifyError::InvalidProofShape); }
            let mut folded = Challenge::zero();
            let mut idx = index;
            for (r, ((rows, path), beta)) in qp.commit_phase_openings.iter().zip(&betas).enumerate() {
                let log_h = log_global_max_height - r; // length of the layer being folded
                folded += ro[log_h];
                let pair = idx >> 1;
                let evals = rows.first().filter(|e| e.len() == 2).ok_or(VerifyError::InvalidProofShape)?;
                if evals[idx & 1] != folded { return Err(VerifyError::CommitPhaseMmcs(qi)); }
                let mp = MerkleProof { index: pair, heights: vec![1 << (log_h - 1)], siblings: path.clone() };
                BFMmcs::<Challenge>::verify_batch(&self.mmcs, 0, rows, &mp, &digest_from(&proof.commit_phase_commits[r]))
                    .map_err(|_| VerifyError::CommitPhaseMmcs(qi))?;
                // fold_row (two_adic_pcs.rs:87-114): interpolate the pair at beta
                let x0 = Val::two_adic_generator(log_h).exp_u64(reverse_bits_len(pair, log_h - 1) as u64);
                let x1 = -x0;
                folded = evals[0] + (*beta - x0) * (evals[1] - evals[0]) * (x1 - x0).inverse();
                idx = pair;
            }
            if folded != proof.final_poly { return Err(VerifyError::FinalPolyMismatch); } // verifier.rs:84-88
        }
        Ok(())
    }
}

/// The Bitcoin-script side (`generate_verify_expr`, fri/src/two_adic_pcs.rs:532-675) builds a script that checks
/// Taproot openings; it has no meaning for a Merkle commitment and is outside this crate's scope.
impl<ChallengerDsl, ManagerAssign> PcsExpr<Challenge, GpuChallenger, ChallengerDsl, ManagerAssign> for GpuTwoAdicFriPcs {
    type DslRep = script_expr::Dsl<Challenge>;
    fn generate_verify_expr(&self, _rounds: Vec<(Self::Commitment, Vec<(Self::Domain, Vec<(Challenge, Vec<Challenge>)>)>)>,
                            _proof: &Self::Proof, _challenger: &mut GpuChallenger, _challenger_dsl: &mut ChallengerDsl)
        -> Result<(ManagerAssign, Vec<Self::DslRep>), Self::Error> {
        unimplemented!("script verifier generation is the reference's CPU/Bitcoin-script subsystem (out of scope)")
    }
}

/// `fri::fold_even_odd` (fri/src/fold_even_odd.rs:20-52) on host vectors: bit-reversed evaluations in, folded out.
pub fn fold_even_odd(poly: Vec<Challenge>, beta: Challenge) -> Vec<Challenge> {
    let h = poly.len() / 2;
    let mut out: Vec<Challenge> = Vec::with_capacity(h);
    CTX.with(|c| {
        let b: Vec<u32> = beta.as_base_slice().iter().map(|&x| monty(x)).collect();
        let rc = unsafe { sys::ts_fri_fold_ext_host(c.0, poly.as_ptr() as *const u32, h, b.as_ptr(), out.as_mut_ptr() as *mut u32) };
        c.check(rc, "ts_fri_fold_ext_host");
    });
    unsafe { out.set_len(h) };
    out
}

// ------------------------------------------------------------------------------------------------ TapTree commitment
/// The device side of `TCS::commit_polys` (basic/src/tcs/mod.rs:238-282): `CompleteTaptree::new_with_scripts` over leaf scripts that
/// share their bit-commitment locking scripts.  The caller passes the script as a template -- the `n + 1` constant byte runs around
/// the `n` pushed integers (`{self.index}` first, then the evaluation limbs in `generate_script`'s order, tcs/mod.rs:197-225) -- and
/// which word of the padded leaf row each later push takes; nothing per leaf is built on the host.
pub struct GpuTapTree {
    raw: *mut sys::ts_taptree,
    pub root: [u8; 32],
    pub n_leaves: usize,
}
impl Drop for GpuTapTree { fn drop(&mut self) { unsafe { sys::ts_taptree_free(self.raw) } } }
impl GpuTapTree {
    /// `matrices`: Montgomery words as p3 stores them; the padded rows (`PolyTCS::padding_matrix`) are built on the device.
    pub fn commit(matrices: &[RowMajorMatrix<BabyBear>], segments: &[Vec<u8>], push_word: &[u32]) -> Self {
        assert_eq!(segments.len(), push_word.len() + 2, "one segment more than pushes");
        CTX.with(|c| {
            let mut devs: Vec<*mut sys::ts_matrix> = Vec::with_capacity(matrices.len());
            for m in matrices {
                let mut d = std::ptr::null_mut();
                let rc = unsafe { sys::ts_matrix_from_host(c.0, m.values.as_ptr() as *const u32, m.height(), m.width(), &mut d) };
                c.check(rc, "ts_matrix_from_host");
                devs.push(d);
            }
            let consts: Vec<*const sys::ts_matrix> = devs.iter().map(|d| *d as *const _).collect();
            let mut rows = std::ptr::null_mut();
            let rc = unsafe { sys::ts_padded_leaf_rows(c.0, consts.as_ptr(), consts.len(), &mut rows) };
            c.check(rc, "ts_padded_leaf_rows");
            let blob: Vec<u8> = segments.concat();
            let mut offs = vec![0usize];
            for s in segments { offs.push(offs.last().unwrap() + s.len()); }
            let (mut root, mut raw) = ([0u8; 32], std::ptr::null_mut());
            let rc = unsafe { sys::ts_taptree_commit(c.0, rows, blob.as_ptr(), offs.as_ptr(), push_word.as_ptr(), push_word.len() + 1,
                                                     root.as_mut_ptr(), &mut raw) };
            let n_leaves = matrices.iter().map(|m| m.height()).max().unwrap_or(0);
            unsafe { sys::ts_matrix_free(rows); for d in devs { sys::ts_matrix_free(d); } }
            c.check(rc, "ts_taptree_commit");
            GpuTapTree { raw, root, n_leaves }
        })
    }
    /// `CompleteTaptree::leaf_indices` (reverse_idx_dict, builder.rs:96-102).
    pub fn leaf_indices(&self) -> Vec<u32> {
        let mut v = vec![0u32; self.n_leaves];
        CTX.with(|c| c.check(unsafe { sys::ts_taptree_leaf_indices(c.0, self.raw, v.as_mut_ptr()) }, "ts_taptree_leaf_indices"));
        v
    }
    /// The `TaprootMerkleBranch` of Merkle leaf `index` (leaf level first) and its TapTree position (tcs/mod.rs:141-146).
    pub fn open(&self, index: usize) -> (Vec<[u8; 32]>, u32) {
        let depth = self.n_leaves.trailing_zeros() as usize;
        let mut path = vec![0u8; 32 * depth.max(1)];
        let mut pos = 0u32;
        CTX.with(|c| c.check(unsafe { sys::ts_taptree_open(c.0, self.raw, index, path.as_mut_ptr(), &mut pos) }, "ts_taptree_open"));
        ((0..depth).map(|l| path[32 * l..32 * l + 32].try_into().unwrap()).collect(), pos)
    }
}
