/*
 * sobolev_b200 -- C ABI of the B200-native kernel-ridge-regression hot path.
 *
 * This is the boundary a maintainer of NKI-CCB/sobolev_alignment binds in place of the third-party
 * `falkon` package that `sobolev_alignment/krr_approx.py:29-34` imports.  The reference is Python, so
 * the binding is a `ctypes.CDLL` (see INTEGRATION.md and sobolev_alignment_b200/_native.py); every
 * entry point below names the reference interface (file:line in /root/reference) it replaces.
 * `[ext]` marks an interface that lives in the external `falkon` dependency and is reached through the
 * cited reference call site.
 *
 * Conventions
 *   - return value 0 = success, anything else = failure; `sa_last_error()` then returns a
 *     thread-local human readable message.  No C++ exception crosses this boundary.
 *   - every pointer is a DEVICE pointer unless the name says host; the caller owns all memory
 *     (the library never allocates device memory and never synchronises the device).
 *   - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it.
 *   - matrices are row-major; `ld*` is the row stride in ELEMENTS.
 *   - "prepared" operands are what `sa_prep` writes: fp16 rows, centred and scaled so that the
 *     euclidean distance between prepared rows is the kernel argument, row stride a multiple of 64,
 *     zero padded; plus fp32 squared norms of the ROUNDED rows in an array whose length is padded
 *     to a multiple of 256 (padding = 0).
 *   - `dp` is the padded number of right-hand-side columns: multiple of 4, 4 <= dp <= 32.
 */
#ifndef SOBOLEV_B200_H_
#define SOBOLEV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kernel families: falkon.kernels.{GaussianKernel,LaplacianKernel,MaternKernel} [ext],
 * constructed at krr_approx.py:67-70,182 */
enum {
  SA_KERNEL_GAUSSIAN = 0, /* exp(-kscale * D^2)                      kscale = 1 / (2 sigma^2)      */
  SA_KERNEL_LAPLACIAN = 1, /* exp(-r),                  r = kscale*D  kscale = 1 / sigma            */
  SA_KERNEL_MATERN15 = 2,  /* (1 + sqrt3 r) exp(-sqrt3 r)                                           */
  SA_KERNEL_MATERN25 = 3   /* (1 + sqrt5 r + 5 r^2/3) exp(-sqrt5 r)                                 */
};

int sa_version(void);
const char* sa_last_error(void);

/* Tuning / diagnostic options, e.g. sa_set_option("SAB_TRSM_CHUNK", "64").  The
 * SAB_* environment variable of the same name is the default, read once at its first use; value NULL
 * restores it.  Not on the reference's path. */
int sa_set_option(const char* name, const char* value);

/* Number of SMs / bytes of dynamic shared memory the fused kernel was configured with for the
 * current device (diagnostics for bench.py). */
int sa_device_info(int* sm_count, int* smem_bytes);

/*
 * Row preparation (falkon `square_norm` + the implicit fp32 operands of `fmmv` [ext]; the input is
 * what `KRRApprox.fit` hands over at krr_approx.py:227 / `transform` at :314):
 *   out[i, j] = fp16( (X[i, j] - shift[j]) * scale[j] * gscale ),  j < p ;  0 for p <= j < ldo
 *   norms[i]  = sum_j float(out[i, j])^2 ;  norms[i] = 0 for n <= i < norms_len
 * `shift` / `scale` may be NULL (0 / 1).  X is never written.
 */
int sa_prep(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, const float* shift,
            const float* scale, float gscale, void* out_f16, int64_t ldo, float* norms,
            int64_t norms_len);
/* Same with f(x) = log10(x + 1) applied to X first (raw counts in): the reference's preprocessing chain
 * (sobolev_alignment.py:851-902: log10(x + 1), StandardScaler, Frobenius gain) is affine after the log, so the
 * caller folds it into shift / scale and no fp32 intermediate is ever written. */
int sa_prep_log(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, const float* shift,
                const float* scale, float gscale, void* out_f16, int64_t ldo, float* norms,
                int64_t norms_len);

/*
 * Input preprocessing of the artificial samples before the fit (SobolevAlignment._memmap_log_processing
 * and _frobenius_normalisation, sobolev_alignment.py:830-902: np.log10(x + 1), StandardScaler, one
 * scalar per data source), as two streaming passes on the device.
 * sa_colstats: sum[j] += sum_i y_ij, sumsq[j] += sum_i y_ij^2 with y_ij = f(X[i, j]) - pivot[j], f = log10(x + 1)
 * when log10p1 != 0, else the identity; pivot (may be NULL = 0) should lie near the column mean (e.g. f of the
 * first row) so that the variance sumsq/n - (sum/n)^2 does not cancel; float64 accumulators, ACCUMULATED INTO
 * (the caller zeroes them; at most 65535 * 4096 rows per call).
 * sa_transform_rows: out[i, j] = (f(X[i, j]) - shift[j]) * scale[j] * gain (shift / scale may be NULL; out may
 * alias X, or be NULL when only the row-norm sum is wanted); rownorm_sum (may be NULL) += sum_i |out[i, :]|_2, the numerator of the mean row norm of the
 * Frobenius normalisation.
 */
int sa_colstats(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, int log10p1,
                const float* pivot, double* sum, double* sumsq);
int sa_transform_rows(void* stream, const float* X, int64_t n, int64_t p, int64_t ldx, int log10p1,
                      const float* shift, const float* scale, float gain, float* out, int64_t ldo,
                      double* rownorm_sum);

/*
 * Fused kernel-matrix x matrix product, K never materialised (falkon `Kernel.mmv` and each half of
 * `Kernel.dmmv` [ext]; reached from `Falkon.fit` krr_approx.py:227 and `Falkon.predict` :314):
 *   out[a x dp] += k(A, B) @ V[b x dp]
 * A [a x p_pad], B [b x p_pad] are prepared operands with norms nA, nB.  V and out are fp32,
 * row-major, dense (row stride dp).  `out` is accumulated into with atomic adds (the caller zeroes
 * it for a plain product).  `symmetric` != 0 asserts A == B so the exact zero distance is used on
 * the diagonal.  `workspace` is caller-owned scratch of at least sa_kmv_workspace_bytes(b, dp)
 * bytes, 128-byte aligned: the library writes the per-column scales of V and its fp16 hi/lo tiles
 * there before the main kernel (same stream); it may be reused by the next call on that stream.
 */
int64_t sa_kmv_workspace_bytes(int64_t b, int dp);
int sa_kmv(void* stream, int kernel, float kscale, const void* A, const float* nA, int64_t a,
           int64_t lda, const void* B, const float* nB, int64_t b, int64_t ldb, int64_t p_pad,
           const float* V, int dp, float* out, int symmetric, void* workspace,
           int64_t workspace_bytes);

/*
 * Optional "stored block" form of one K_nM^T (K_nM z) product (falkon keeps K_nM when it fits --
 * `FalkonOptions.never_store_kernel = False`, krr_model_selection.py:56 -- here it is a transient ROW BLOCK):
 *   sa_kmv_store  as sa_kmv, and the kernel block k(A, B) is also written to `kstore` as fp16 hi/lo pairs in
 *                 packed 128 x 128 tiles (4 bytes per entry; sa_kstore_bytes(a, b) bytes, 128-byte aligned)
 *   sa_ktv        out[b x dp] += k(A, B)^T t  streamed from that store (HBM-bound: the block is read once)
 *                 instead of forming K a second time; t [a x dp] fp32; workspace sa_ktv_workspace_bytes(a).
 * The default path never materialises K (sa_kmv twice per product pair).
 */
int64_t sa_kstore_bytes(int64_t a, int64_t b);
int sa_kmv_store(void* stream, int kernel, float kscale, const void* A, const float* nA, int64_t a,
                 int64_t lda, const void* B, const float* nB, int64_t b, int64_t ldb, int64_t p_pad,
                 const float* V, int dp, float* out, void* workspace, int64_t workspace_bytes, void* kstore,
                 int64_t kstore_bytes);
int64_t sa_ktv_workspace_bytes(int64_t a);
int sa_ktv(void* stream, const void* kstore, int64_t a, int64_t b, const float* t, int dp, float* out,
           void* workspace, int64_t workspace_bytes);

/*
 * Dense kernel matrix (falkon `Kernel.__call__` [ext]; called as `kernel_(A, B)` at
 * sobolev_alignment.py:950,955-958 and tests/test_krr_approx.py:255, and for K_MM inside
 * `Falkon.fit`):  out[a x b] = k(A, B), fp32, row stride ldo.
 */
int sa_kdense(void* stream, int kernel, float kscale, const void* A, const float* nA, int64_t a,
              int64_t lda, const void* B, const float* nB, int64_t b, int64_t ldb, int64_t p_pad,
              float* out, int64_t ldo, int symmetric);

/*
 * Nystrom preconditioner (falkon `FalkonPreconditioner.init`: potrf / lauum / mul_triang [ext],
 * reached from krr_approx.py:227).  All matrices fp32 row-major, LOWER triangle significant;
 * T = L^T and A = LA^T are the upper factors of the published algorithm.
 *
 * sa_potrf: in-place blocked Cholesky A = L L^T (lower).  `invdiag` receives the inverses of the
 * 128x128 diagonal blocks of L ( ceil(M/128) blocks of 128*128 floats ); they are what the
 * triangular solves use.  `info` (device int) is set to 1 + index of the first non-positive pivot,
 * or left 0.  The trailing updates run on the tensor cores with fp16 hi/lo operand pairs staged in
 * `workspace` (caller-owned, 256-byte aligned, at least sa_linalg_workspace_bytes(M) bytes; contents
 * are scratch).  Elements above the 128-block diagonal are scratch as well.
 */
int64_t sa_linalg_workspace_bytes(int64_t M);
int sa_potrf(void* stream, float* A, int64_t M, int64_t lda, float* invdiag, int* info, void* workspace,
             int64_t workspace_bytes);

/*
 * The same factorisation in steps, for the row-sharded (multi-GPU) solver: the outer panels of
 * sa_potrf_outer_blocks() block columns are dealt out over the ranks, the owner factors a panel, the panel
 * travels by NCCL broadcast (outside this ABI) and every rank applies it to the panels it owns.
 *   sa_potrf_begin   once: clears `info`, derives the operand scale of the tensor-core updates
 *   sa_potrf_panel   factor block columns [k0, kend) (all rows below) in place; their earlier updates
 *                    must have been applied
 *   sa_potrf_split   fp16 hi/lo operands of a finished panel; `panel` = its rows BELOW its diagonal
 *                    blocks, [rows x W] fp32 with row stride ldp (a slice of A, or a received copy)
 *   sa_potrf_update  C -= P P^T restricted to block columns [c0, c1) (rows from block c0 down) of the
 *                    trailing matrix that starts at block kend, P = the panel last split
 * sa_potrf == begin; for each panel: panel, split, update(kend, kend, M/128).
 */
int sa_potrf_outer_blocks(void);
int sa_potrf_begin(void* stream, const float* A, int64_t M, int64_t lda, int* info, void* workspace,
                   int64_t workspace_bytes);
int sa_potrf_panel(void* stream, float* A, int64_t M, int64_t lda, float* invdiag, int* info, int k0,
                   int kend);
int sa_potrf_split(void* stream, int64_t M, const float* panel, int64_t ldp, int64_t rows, int W,
                   void* workspace, int64_t workspace_bytes);
int sa_potrf_update(void* stream, float* A, int64_t M, int64_t lda, int kend, int c0, int c1,
                    const float* panel, int64_t ldp, int W, void* workspace, int64_t workspace_bytes);
/* sa_potrf_update for every outer panel (nb blocks wide) a rank owns in one call: panels first, first + stride, ...
 * except `skip` (-1: none). */
int sa_potrf_update_cyclic(void* stream, float* A, int64_t M, int64_t lda, int kend, int nb, int first, int stride,
                           int skip, const float* panel, int64_t ldp, int W, void* workspace,
                           int64_t workspace_bytes);

/* G(lower) = alpha * L^T L + beta * I   (falkon `lauum` + `mul_triang` + add-diagonal [ext]);
 * same workspace contract as sa_potrf. */
int sa_ltl(void* stream, const float* L, int64_t M, int64_t ldl, float* G, int64_t ldg, float alpha,
           float beta, void* workspace, int64_t workspace_bytes);
/* Multi-GPU form: rank `part` of `nparts` accumulates its share of the 512-row slabs of L into G (zeroed by the
 * caller, no beta term); the caller sums the parts (all-reduce) and adds beta with sa_add_diag.  The slabs are
 * dealt out from the largest down, so the parts are balanced. */
int sa_ltl_part(void* stream, const float* L, int64_t M, int64_t ldl, float* G, int64_t ldg, float alpha,
                float beta, void* workspace, int64_t workspace_bytes, int part, int nparts);

/* A[i, i] += value */
int sa_add_diag(void* stream, float* A, int64_t M, int64_t lda, float value);

/*
 * Triangular solves with a Cholesky factor from sa_potrf (cuBLAS `trsm` inside falkon's
 * `invT / invTt / invA / invAt` [ext], 4 per CG iteration of the fit at krr_approx.py:227).
 *
 * The factor is constant over the ~90 solves of a fit, so it is PACKED once: every strictly-lower
 * 128 x 128 tile, the inverted diagonal blocks and the pre-multiplied sub-diagonal tiles
 * inv(L_cc) L(c, c-1) / L(c+1, c) inv(L_cc) as fp16 pairs (x 2^s = hi + lo: the bytes of fp32, ~22 bits)
 * in the shared-memory image of the solve kernel.  `packed` is caller-owned, sa_trsm_packed_bytes(M) bytes
 * (about HALF of the fp32 square, which may be released afterwards), 128-byte aligned;
 * `workspace` of sa_trsm_pack holds sa_trsm_pack_workspace_bytes(M) bytes (fp32 staging of the
 * pre-multiplied tiles).
 */
int64_t sa_trsm_packed_bytes(int64_t M);
int64_t sa_trsm_pack_workspace_bytes(int64_t M);
int sa_trsm_pack(void* stream, const float* L, int64_t M, int64_t ldl, const float* invdiag, void* packed,
                 void* workspace, int64_t workspace_bytes);

/*
 * B[M x dp] <- L^-1 B (transpose = 0) or L^-T B (transpose = 1) with a packed factor.  One persistent
 * launch per solve: block rows are cut into chunks of dependency tiles that CTAs draw from a ticket, the
 * block-to-block chain runs through flags in `workspace` (caller-owned scratch of at least
 * sa_trsm_workspace_bytes(M) bytes, 128-byte aligned; the flags are zeroed by the library on the same
 * stream).  Sums are formed in a fixed order: results are bit-reproducible.
 */
int64_t sa_trsm_workspace_bytes(int64_t M);
int sa_trsm_solve(void* stream, const void* packed, int64_t M, float* B, int dp, int transpose,
                  void* workspace, int64_t workspace_bytes);

/*
 * Two dependent solves in one launch (the preconditioner is always applied as such a pair: A^-1 then
 * T^-1, and T^-T then A^-T, `FalkonPreconditioner.apply / apply_t` [ext]):
 *   Xa <- op(La)^-1 Xa                       (in place)
 *   Xb <- op(Lb)^-1 (Xa + lambda * V)        (V may be NULL)
 * op = identity (transpose = 0) or transposition (1) for both.  The second chain takes block c as soon
 * as the first has published it, so the pair costs one chain latency.  Xa != Xb; same workspace contract.
 */
int sa_trsm_solve_pair(void* stream, const void* packed_a, const void* packed_b, int64_t M, float* Xa,
                       float* Xb, const float* V, float lambda, int dp, int transpose, void* workspace,
                       int64_t workspace_bytes);

/*
 * Conjugate-gradient vector kernels (falkon `ConjugateGradient.solve` [ext]); all M x dp fp32.
 * sa_coldot:   out[c] = sum_i P[i, c] * Q[i, c]                    (out is overwritten)
 * sa_cg_update: alpha_c = rs_old[c] / (pAp[c] + eps);  X += P alpha;  R -= AP alpha (if AP != NULL);
 *              rs_new[c] = sum_i R[i, c]^2 (if R updated)
 * sa_cg_direction: P = R + rs_new / (rs_old + eps) P
 * sa_axpby:    Y = a * X + b * Y   (elementwise, len floats)
 * Column sums are deterministic (per-block partials in `workspace`, folded in index order by the block
 * that finishes last; no float atomics), so replicated solver state stays bit-identical across ranks.
 * `workspace`: sa_cg_workspace_bytes() bytes, zeroed ONCE by the caller before its first use.
 * Convergence is decided on the device: when `hist` != NULL the sums are also stored to hist[0..dp)
 * (one row of the residual history); when `stop` != NULL, *stop is raised once max_c sqrt(sum_c) < tol,
 * and sa_cg_update / sa_cg_direction become no-ops while *stop != 0 (the host reads the flag late,
 * without synchronising every iteration -- falkon reads the residual every iteration).
 */
int64_t sa_cg_workspace_bytes(void);
int sa_coldot(void* stream, const float* P, const float* Q, int64_t M, int dp, float* out, void* workspace,
              float* hist, int* stop, float tol);
int sa_cg_update(void* stream, int64_t M, int dp, const float* P, const float* AP, float* X,
                 float* R, const float* rs_old, const float* pAp, float eps, float* rs_new, void* workspace,
                 float* hist, int* stop, float tol);
int sa_cg_direction(void* stream, int64_t M, int dp, const float* R, float* P, const float* rs_new,
                    const float* rs_old, float eps, const int* stop);
int sa_axpby(void* stream, int64_t len, float a, const float* X, float b, float* Y);

#ifdef __cplusplus
}
#endif
#endif /* SOBOLEV_B200_H_ */
