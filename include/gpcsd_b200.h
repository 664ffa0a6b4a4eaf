/*
 * gpcsd_b200.h -- C ABI of the B200-native GPCSD hot path (libgpcsd_b200.so).
 *
 * The reference (natalieklein/gpcsd) is pure Python with no FFI; its boundary for this path is the
 * Python object API (GPCSD1D/GPCSD2D.loglik / fit.obj_fun / predict and the covariance helpers).  The
 * entry points below are what a ctypes binding inside those methods would call; each one names the
 * reference code it replaces (paths relative to src/gpcsd/ of the reference).  INTEGRATION.md shows
 * the binding.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  All matrices are FP64, ROW-MAJOR, with an explicit
 *    leading dimension `ld*` counted in doubles.  Device pointers unless the parameter name starts
 *    with `h_` (host).  Leading dimensions and batch strides of GEMM operands must be EVEN (16-byte
 *    rows) and base pointers 16-byte aligned; the Python host layer allocates that way.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *    performs no hidden allocation (workspaces are caller-provided; sizes from the *_ws_* queries).
 *  - Return value: 0 on success, non-zero on error; gpcsd_last_error() returns a thread-local
 *    message.  No C++ exception crosses the ABI.
 *  - LFP layout (reference: `self.lfp`, shape (nx, nt, ntrials), C order => trial index fastest,
 *    gpcsd1d.py:125,255): device array Y[nx][nt][ldn], ldn >= ntrials, ldn even.
 */
#ifndef GPCSD_B200_H
#define GPCSD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPCSD_B200_ABI_VERSION 1

#define GPCSD_KIND_SE 0     /* GPCSDTemporalCovSE      covariances.py:240-271 */
#define GPCSD_KIND_MATERN 1 /* GPCSDTemporalCovMatern  covariances.py:274-305 */

int gpcsd_abi_version(void);
const char* gpcsd_last_error(void);
/* number of SMs of the current device (grid sizing is derived from it) */
int gpcsd_num_sms(void);

/* ---------------------------------------------------------------------------------------------
 * FP64 tensor-core (DMMA) strided-batched GEMM:  C_b = A_b (M x K) * op(B_b),  b = 0..batch-1
 *   transB == 0 : B_b is K x N row-major (N contiguous)       -- replaces np.dot / np.matmul
 *   transB == 1 : B_b is N x K row-major (K contiguous), C = A B^T
 * Replaces every dense product on the path: Qs^T Y_r / (.) Qt of loglik (gpcsd1d.py:125,
 * gpcsd2d.py:148), A Kg / (A Kg) A^T of compKphi (covariances.py:90,95,223,231), the cross-covariance
 * products (covariances.py:71,201) and mykron(..).T @ invy of predict (gpcsd1d.py:279,283).
 * ------------------------------------------------------------------------------------------- */
int gpcsd_dgemm(int transB, int M, int N, int K,
                const double* A, long lda, long strideA,
                const double* B, long ldb, long strideB,
                double* C, long ldc, long strideC,
                int batch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Kronecker projection + quadratic form (hot loop gpcsd1d.py:124-126 / gpcsd2d.py:147-149):
 * for every spatial eigen-index i (batch):  A_i = Qt^T Z_i  (nt x ntrials), Z = Qs^T Y already formed
 * by gpcsd_dgemm;  epilogue  B_i = A_i * rD[i][:],  quad += sum A_i .* B_i,  bsq += sum B_i .* B_i.
 *   QtT    : nt x nt, row-major Q^T  (== column-major eigenvector matrix as cuSOLVER returns it)
 *   Z      : [nx][nt][ldn]            rD : [nx][ldrd] reciprocal of D_ij = ls_i lt_j + sig2n(_i)
 *   Bout   : [nx][nt][ldn]  (= (Qs^T Y_r Qt) / D for all trials; input of the gradient and of predict)
 *   partials: workspace of gpcsd_project_quad_ws_doubles(...) doubles; out2[0] = quad, out2[1] = bsq
 * ------------------------------------------------------------------------------------------- */
long gpcsd_project_quad_ws_doubles(int nx, int nt, int ntrials);
int gpcsd_project_quad(int nx, int nt, int ntrials,
                       const double* QtT, long ldq,
                       const double* Z, long ldn,
                       const double* rD, long ldrd,
                       double* Bout, double* partials, double* out2, void* stream);

/* The same contraction restricted to one block of a block-diagonal temporal eigenbasis (see gpcsd_centro_fold): AT is the
 * m x m transposed eigenvector block, Z / Bout / rD point at the block's first time row / eigen-index inside the parent
 * [nx][nt][ldn] / [nx][ldrd] arrays, bstride = nt*ldn.  Workspace: gpcsd_project_quad_ws_doubles(nx, m, ntrials). */
int gpcsd_project_quad_strided(int nx, int m, int ntrials, const double* AT, long lda, const double* Z, long ldn, long bstride,
                               const double* rD, long ldrd, double* Bout, double* partials, double* out2, void* stream);

/* Restart-batched form: R hyperparameter vectors x nx spatial eigen-indices in ONE launch (the trial loop of loglik for every
 * restart of fit(), gpcsd1d.py:193-211 x 124-126).  Restart r uses the m x m block AT + r*strideA; Z / Bout / rD of consecutive
 * restarts follow each other (Z_r = Z + r*nx*bstride, rD_r = rD + r*nx*ldrd); out2[r*out_stride + {0, 1}] = (quad, bsq). */
long gpcsd_project_quad_batched_ws_doubles(int R, int nx, int m, int ntrials);
int gpcsd_project_quad_batched(int R, int nx, int m, int ntrials, const double* AT, long lda, long strideA, const double* Z, long ldn,
                               long bstride, const double* rD, long ldrd, double* Bout, double* partials, double* out2,
                               long out_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Segment-weighted symmetric rank-k update (gradient of the quadratic form; replaces the autograd
 * tape of gpcsd1d.py:211 / gpcsd2d.py:250 over the trial loop):
 *     C (M x M) = sum_{seg=0}^{nseg-1} w[seg] * X_seg X_seg^T,   X_seg[m][k] = X[m*row_stride + seg*seg_stride + k],
 *     k = 0..seglen-1.   w == NULL means all ones.
 *   Mt = sum_i ls_i B_i^T-contraction : M = nt, row_stride = ldn,     nseg = nx, seg_stride = nt*ldn, w = ls
 *   Ms = sum_j lt_j (...)             : M = nx, row_stride = nt*ldn,  nseg = nt, seg_stride = ldn,    w = lt
 *   Ns = sum_j (...)                  : as Ms with w = NULL (per-electrode-noise mode only)
 * Split-K over (seg, k) with per-CTA partial tiles in `ws` and a fixed-order second pass, so the result
 * is deterministic (no atomics).  Both triangles of C are written.
 * ------------------------------------------------------------------------------------------- */
long gpcsd_wsyrk_ws_doubles(int M, int nseg, int seglen);
int gpcsd_wsyrk(int M, int nseg, int seglen,
                const double* X, long row_stride, long seg_stride,
                const double* w, double* C, long ldc, double* ws, void* stream);

/* Restart-batched SYRK (all restarts of a multi-start batch in one launch): for r < R,
 *   Cw_r = sum_seg w_r[seg] X_r,seg X_r,seg^T   and, when Cp != NULL, Cp_r = sum_seg X_r,seg X_r,seg^T,
 * X_r = X + r*strideX, w_r = w + r*strideW, C_r = C + r*strideC.  ws: gpcsd_wsyrk_batched_ws_doubles(R, M, nseg, seglen, pair). */
long gpcsd_wsyrk_batched_ws_doubles(int R, int M, int nseg, int seglen, int pair);
int gpcsd_wsyrk_batched(int R, int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, long strideX,
                        const double* w, long strideW, double* Cw, double* Cp, long ldc, long strideC, double* ws, void* stream);
/* Cw = sum_seg w[seg] X_seg X_seg^T and Cp = sum_seg X_seg X_seg^T in ONE pass over X when M <= 32 (the Ms / Ns pair of the
 * per-electrode-noise gradient: both read Bm); two passes above.  ws: 2 * gpcsd_wsyrk_ws_doubles(M, nseg, seglen) doubles. */
int gpcsd_wsyrk_pair(int M, int nseg, int seglen, const double* X, long row_stride, long seg_stride, const double* w,
                     double* Cw, double* Cp, long ldc, double* ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Eigendecomposition of the small factors (np.linalg.eigh in comp_eig_D, utility_functions.py:58-59).
 * cuSOLVER syevd on a copy; returns eigenvalues ascending in W and Q^T row-major (== column-major Q)
 * in QT (n x n, leading dimension ldq).  `info` is a device int (0 = converged).
 * ------------------------------------------------------------------------------------------- */
long gpcsd_eigh_ws_doubles(int n, long ldq);
int gpcsd_eigh(int n, const double* K, long ldk, double* QT, long ldq, double* W,
               double* ws, long ws_doubles, int* info, void* stream);

/* Batched variant for small orders (cusolverDnXsyevBatched): `batch` symmetric matrices [n][ld] stacked with stride n*ld
 * are overwritten by their eigenvectors stored as ROWS (Q^T row-major); W: [batch][n] ascending; info: [batch] device ints.
 * On B200 matrices of order <= 128 with batch >= 2 stay inside one CTA each: 0.78 ms for two 128 x 128 problems against
 * 2.1 ms for one syevd(128). */
long gpcsd_eigh_batched_ws_bytes(int n, long ld, int batch);
int gpcsd_eigh_batched(int n, int batch, double* A, long ld, double* W, void* ws, long ws_bytes, int* info, void* stream);

/* In-house symmetric eigensolver for orders 3..256 (gpcsd_eig.cu), one 8-CTA thread-block cluster per matrix, batched:
 * replaces np.linalg.eigh of comp_eig_D (utility_functions.py:58-59) where cuSOLVER syevd is latency-bound.
 * M[nmat][n][ldm] symmetric (not modified) -> QT[nmat][n][ldq] (rows = eigenvectors), W[nmat][n] ascending.
 * info[nmat] (device int array, may be NULL): 0 = ok, 1 = non-finite input (outputs are NaN; numpy.linalg.eigh raises
 * LinAlgError in that case, the Python layer does the same). */
long gpcsd_eigh_dc_ws_doubles(int n, long ldq, int nmat);
int gpcsd_eigh_dc(int n, int nmat, const double* M, long ldm, double* QT, long ldq, double* W, double* ws, long ws_doubles,
                  int* info, void* stream);

/* Its three stages, exported for tests and reuse:
 * (1) Householder tridiagonalisation M = H T H^T (matrix resident in distributed shared memory): d[nmat][n], e[nmat][n]
 *     (e[k] = T[k+1][k]), V[nmat][n][ldv] (row k = reflector k, implicit 1 at column k+1), tau[nmat][n];
 * (2) divide-and-conquer eigen-decomposition of the tridiagonal matrices (Cuppen / Gu-Eisenstat): W ascending, XT rows =
 *     eigenvectors of T; ws = 2*nmat*n*ldx doubles;
 * (3) back-transformation of eigenvectors stored as rows, in place. */
int gpcsd_tridiag(int n, int nmat, const double* M, long ldm, double* d, double* e, double* V, long ldv, double* tau,
                  void* stream);
long gpcsd_tridiag_eig_ws_doubles(int n, long ldx, int nmat);
int gpcsd_tridiag_eig(int n, int nmat, const double* d, const double* e, double* W, double* XT, long ldx, double* ws,
                      long ws_doubles, int* info, void* stream);
int gpcsd_backtransform(int n, int nmat, const double* V, long ldv, const double* tau, double* XT, long ldx, void* stream);

/* Exact centrosymmetric split of a symmetric Toeplitz (more generally J K J = K) matrix -- every stationary
 * temporal kernel of covariances.py:257-305 on a uniform time grid -- into two independent half-size
 * eigenproblems (Cantoni & Butler 1976), so the np.linalg.eigh(Kt) of utility_functions.py:58 costs two
 * concurrent syevd of order n/2:  S = K11 + K12 J ((n/2 + n%2)^2, bordered for odd n),  A = K11 - K12 J ((n/2)^2).
 * gpcsd_centro_assemble rebuilds QT (rows = eigenvectors of K) and W = [Ws, Wa] (unsorted). */
int gpcsd_centro_split(int n, const double* K, long ldk, double* S, long lds, double* A, long lda, void* stream);
int gpcsd_centro_assemble(int n, const double* UsT, long lds, const double* Ws, const double* UaT, long lda,
                          const double* Wa, double* QT, long ldq, double* W, void* stream);

/* Centrosymmetric fold of the time axis of a trial block X[nblk][n][rowlen] (trials contiguous, rowlen even):
 * Xf[b][j] = (X[b][j] + X[b][n-1-j])/sqrt2, Xf[b][ms+j] = (X[b][j] - X[b][n-1-j])/sqrt2 (j < n/2; middle row kept for odd n).
 * In this basis the eigenvector matrix assembled by gpcsd_centro_assemble is block diagonal (Us^T, Ua^T), so the trial loop
 * of loglik (gpcsd1d.py:124-126) and the temporal SYRK of its gradient cost half the flops. */
int gpcsd_centro_fold(int nblk, int n, long rowlen, const double* X, double* Xf, void* stream);
/* Its inverse (the fold is orthogonal): folded time rows back to the time axis.  predict (gpcsd1d.py:279-291) at t* = t applies
 * Kt*_k Qt as two half-order products in the folded basis and unfolds the result. */
int gpcsd_centro_unfold(int nblk, int n, long rowlen, const double* Xf, double* X, void* stream);

/* Same split for any fixed-point-free involution pi with K[pi(i)][pi(j)] == K[i][j] (n even): device int arrays
 * ra[n/2] (representatives) and rb[n/2] = pi(ra).  Used for the spatial factor of geometries that are invariant under the
 * point reflection about the centre of the integration box (Neuropixels checkerboard, covariances.py:204-232). */
int gpcsd_pairsym_split(int n, const double* K, long ldk, const int* ra, const int* rb, double* S, long lds, double* A,
                        long lda, void* stream);
int gpcsd_pairsym_assemble(int n, const int* ra, const int* rb, const double* UsT, long lds, const double* Ws,
                           const double* UaT, long lda, const double* Wa, double* QT, long ldq, double* W, void* stream);

/* Channel-axis analogue of gpcsd_centro_fold for that involution: X[n][rowlen] -> Xf[c] = (X[ra[c]] + X[rb[c]])/sqrt2,
 * Xf[n/2 + c] = (X[ra[c]] - X[rb[c]])/sqrt2.  Qs (as assembled above) is block diagonal in this basis, so Qs^T Y (the spatial
 * half of the trial loop gpcsd2d.py:147-149) and the spatial SYRK of the gradient cost half the flops. */
int gpcsd_pairsym_fold(int n, const int* ra, const int* rb, long rowlen, const double* X, double* Xf, void* stream);

/* ---------------------------------------------------------------------------------------------
 * D and its reductions (utility_functions.py:54-63; gpcsd1d.py:122):
 *   D_ij = ls_i * lt_j + s_i  (s = sig2n[0] if n_sig2n == 1 else sig2n[i], i = ASCENDING spatial eigen-index)
 *   rD[i][j] = 1 / D_ij ;  sums[0] = sum log D ;  sums[1] = sum 1/D
 *   rowA[i] = sum_j lt_j / D_ij ; rowC[i] = sum_j 1 / D_ij ; rowL[i] = sum_j log D_ij ; colB[j] = sum_i ls_i / D_ij
 * ------------------------------------------------------------------------------------------- */
int gpcsd_eig_D(int nx, int nt, const double* ls, const double* lt, const double* sig2n, int n_sig2n,
                double* rD, long ldrd, double* sums2, double* rowA, double* rowC, double* rowL, double* colB,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Forward-model quadrature weights  A = gl_w (.) b_fwd  and dA/dR.
 *   1-D: b_fwd_1d(gl_x - x, R) forward_models.py:9-17 inside compKphi_1d covariances.py:86-88
 *   2-D: b_fwd_2d(.., R, eps, w=delta_w) forward_models.py:42-54 inside compKphi_2d covariances.py:220-221;
 *        quadrature node g = (g1, g2) of the x1-major product grid (utility_functions.py:22), g = g1*ngl2 + g2
 * A, dA: [npts][ldg]; dA may be NULL.
 * ------------------------------------------------------------------------------------------- */
int gpcsd_fwd_weights_1d(int npts, const double* x, int G, const double* gl_x, const double* gl_w,
                         double R, double* A, double* dA, long ldg, void* stream);
int gpcsd_fwd_weights_2d(int npts, const double* pts /* [npts][2] */, int ngl1, int ngl2,
                         const double* gl_x1, const double* gl_w1, const double* gl_x2, const double* gl_w2,
                         double R, double eps, double* A, double* dA, long ldg, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Squared-exponential factor matrices:  out[i][j] = scale * exp(-0.5 ((a_i - b_j)/ell)^2)  (deriv == 0)
 *                                        ... * (a_i - b_j)^2 / ell^3                        (deriv == 1)
 * CSD kernel on the quadrature grid (covariances.py:89, 216 -- the 2-D kernel is the Kronecker product
 * of two such factors on the product grid), CSD-CSD kernel (covariances.py:56,186), gl-to-z kernels of
 * the cross-covariances (covariances.py:67,198).
 * ------------------------------------------------------------------------------------------- */
int gpcsd_se_matrix(int na, const double* a, int nb, const double* b, double ell, double scale, int deriv,
                    double* out, long ld, void* stream);
/* 2-D gl-to-z kernel, stored transposed: out[z][g] = exp(-.5((g1-z1)/ell1)^2) * exp(-.5((g2-z2)/ell2)^2),
 * g = g1*ngl2 + g2  (covariances.py:198; ld >= ngl1*ngl2) */
int gpcsd_se_grid_to_pts(int ngl1, int ngl2, const double* gl_x1, const double* gl_x2,
                         int nz, const double* z /* [nz][2] */, double ell1, double ell2,
                         double* out, long ld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Temporal covariance  Kt[i][j] = sum_k sigma2_k f_k(t_i - tp_j)  (compute_Kt covariances.py:257-271,
 * 291-305 summed as in gpcsd1d.py:118-120).  h_kind/h_ell/h_sigma2 are HOST arrays of length ntc (<= 8).
 * ------------------------------------------------------------------------------------------- */
int gpcsd_kt_build(int nt_rows, const double* t, int nt_cols, const double* tp, int ntc,
                   const int* h_kind, const double* h_ell, const double* h_sigma2,
                   double* Kt, long ld, void* stream);
/* out[2k] = <G, dKt_k/d ell_k>, out[2k+1] = <G, dKt_k/d sigma2_k>;  G: nt x nt (ldg);
 * ws: gpcsd_kt_grad_ws_doubles(nt, ntc) doubles. */
long gpcsd_kt_grad_ws_doubles(int nt, int ntc);
int gpcsd_kt_grad(int nt, const double* t, int ntc, const int* h_kind, const double* h_ell,
                  const double* h_sigma2, const double* G, long ldg, double* ws, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Eigen-basis gradient cores (closed form of the reverse pass):
 *   X[i][i'] = 0.5*M[i][i'] (+ 0.5*(s_i-s_i')/(l_i-l_i')*Nmat[i][i'] if Nmat != NULL), i != i'
 *   X[i][i]  = -0.5*ntrials_total*rowsum[i] + 0.5*M[i][i]
 * `scale_data` multiplies the data terms (M, Nmat) and `scale_det` the log-det term, so that trial-sharded
 * ranks can form partial gradients (data terms local; log-det term once).
 * ------------------------------------------------------------------------------------------- */
int gpcsd_grad_core(int n, const double* Mmat, long ldm, const double* Nmat, long ldnm,
                    const double* lam, const double* s, const double* rowsum, double ntrials_total,
                    double scale_data, double scale_det, double* X, long ldx, void* stream);

/* small helpers (all row-major, FP64) */
int gpcsd_add_diag(int n, double* K, long ld, double v, void* stream);               /* + JITTER*I  gpcsd1d.py:117 */
int gpcsd_transpose(int rows, int cols, const double* in, long ldi, double* out, long ldo, void* stream);
long gpcsd_dot_ws_doubles(long n);
int gpcsd_dot(int rows, int cols, const double* X, long ldx, const double* Y, long ldy, double* ws, double* out, void* stream);
int gpcsd_sum_arrays(long n, int narr, const double* const* h_in, double* out, void* stream);   /* csd += csd_tmp gpcsd1d.py:281 */
int gpcsd_sum_vec(long n, const double* in, double* out, void* stream);

/* =============================================================================================
 * gpcsd_plan: ONE call per loglik + gradient evaluation, batched over R hyperparameter vectors.
 *
 * Replaces obj_fun + autograd.grad(obj_fun) inside the restart loop of fit() (gpcsd1d.py:153-211, gpcsd2d.py:177-260): the
 * covariance build, both eigendecompositions, the Kronecker projection of every trial, the gradient SYRKs and the closed-form
 * reverse pass for R restarts are enqueued by one native driver; the hyperparameters live in device memory, so the launch
 * sequence is captured once per (R, mode) into a CUDA graph and replayed.  theta and the gradient are in NATURAL units in the
 * order R, ell (1-D) or ell1, ell2 (2-D), (ell_t, sigma2_t) per temporal kernel, sig2n (1 or nx entries) -- the order of the
 * reference's tparams (gpcsd1d.py:160-174) without the log transform, which stays in the host layer with the priors.
 * ============================================================================================= */
/* Geometry: x [nx] (1-D) or [nx][2] (2-D), t [nt]; Gauss-Legendre nodes / weights on the integration box (covariances.py:22-27,
 * 114-124; 1-D: g1/w1 of even length G1, G2 = 0; 2-D: product grid G1 x G2 with G2 even, pad with zero-weight nodes);
 * temporal kernel kinds (GPCSD_KIND_*); n_sig2n = 1 or nx; jitter (gpcsd1d.py:17 / gpcsd2d.py:16); eps (2-D).
 * t_uniform != 0 enables the centrosymmetric split of Kt (even nt >= 32) and, from nt >= fold_min_nt, the folded time basis;
 * h_ra / h_rb (nx/2 ints each, or NULL): site pairing of a reflection-symmetric geometry (used with scalar noise, nx >= 64).
 * max_restarts bounds R.  All h_* arrays are HOST pointers and are copied. */
int gpcsd_plan_create(void** plan, int dim, int nx, int nt, const double* h_x, const double* h_t, int G1, const double* h_g1,
                      const double* h_w1, int G2, const double* h_g2, const double* h_w2, int ntc, const int* h_kinds,
                      int n_sig2n, double jitter, double eps, int t_uniform, int fold_min_nt, const int* h_ra, const int* h_rb,
                      int max_restarts);
int gpcsd_plan_destroy(void* plan);
int gpcsd_plan_num_params(void* plan);                          /* P */
/* Bind this rank's trial slab Y[nx][nt][ldn] (device, caller-owned, zero padded; self.lfp of gpcsd1d.py:125) and a caller
 * allocated workspace of gpcsd_plan_ws_bytes(plan, ldn, ntrials_local) bytes.  ntrials_total: trials of the whole model (all
 * ranks); det_fraction: share of the trial-independent log-det terms this rank contributes (1 / world). */
long gpcsd_plan_ws_bytes(void* plan, long ldn, int ntrials_local);
int gpcsd_plan_set_lfp(void* plan, const double* Y, long ldn, int ntrials_local, double ntrials_total, double det_fraction,
                       void* ws, long ws_bytes);
int gpcsd_plan_touch_lfp(void* plan);                           /* the bound Y buffer was overwritten in place */
int gpcsd_plan_set_graph(void* plan, int enable);               /* CUDA-graph replay on (default) / off */
/* h_theta [R][P] host -> h_out [R][P+4] host: loglik, d loglik / d theta (P entries; zeros when want_grad == 0), eigensolver
 * flag (0 = ok; numpy raises LinAlgError otherwise), and a hyperparameter checksum pair (c*f, c^2*f) for rank-consistency
 * checks after an all-reduce.  gpcsd_plan_loglik_grad = gpcsd_plan_enqueue + gpcsd_plan_finish; trial-sharded callers
 * all-reduce gpcsd_plan_device_result() ([R][P+4], on `stream`) between the two.  One evaluation in flight per plan.
 * When calls for DIFFERENT plans overlap in time (several models evaluated from several host threads) gpcsd_plan_loglik_grad
 * issues the evaluation as two launches -- covariances + eigendecompositions, then the kernels that fill the GPU under a
 * process-wide token -- so that one model's latency-bound prologue runs underneath another model's GEMMs; results are
 * bit-identical to the one-launch form (GPCSD_GEMM_TOKEN, GPCSD_GEMM_RESERVE). */
int gpcsd_plan_loglik_grad(void* plan, int R, const double* h_theta, int want_grad, double* h_out, void* stream);
int gpcsd_plan_enqueue(void* plan, int R, const double* h_theta, int want_grad, void* stream);
int gpcsd_plan_finish(void* plan, int R, double* h_out, void* stream);
double* gpcsd_plan_device_result(void* plan);
double* gpcsd_plan_device_theta(void* plan);
long gpcsd_plan_last_launches(void* plan);                      /* kernels launched by the last non-replayed evaluation */
/* Trial-sharded models on one node: the per-evaluation all-reduce of the [R][P+4] result (the sum over `trial` of
 * gpcsd1d.py:124-126 split over ranks) through a host-mapped mailbox.  h_shared: ONE zero-initialised POSIX shared-memory
 * segment of gpcsd_plan_mailbox_bytes(plan, world) bytes mapped by every rank.  After gpcsd_plan_set_mailbox the last kernel of
 * every evaluation writes this rank's result and a sequence flag into its slot, and gpcsd_plan_finish sums all slots in rank
 * order on the host: no collective kernel, no device-side waiting, no device->host copy.  Every rank must issue the same
 * sequence of evaluations. */
long gpcsd_plan_mailbox_bytes(void* plan, int world);
int gpcsd_plan_set_mailbox(void* plan, void* h_shared, long bytes, int world, int rank);
/* Kernel-level entry: one evaluation with CALLER-SUPPLIED eigen-factors instead of the eigensolvers:
 * device arrays QsT [nx][even(nx)], ls [nx], QtT [nt][even(nt)], lt [nt], rows = eigenvectors (the columns np.linalg.eigh
 * returns in comp_eig_D, utility_functions.py:58-59). */
int gpcsd_plan_loglik_grad_factors(void* plan, const double* h_theta, const double* QsT, const double* ls, const double* QtT,
                                   const double* lt, int want_grad, double* h_out, void* stream);

/* =============================================================================================
 * Callers either side of the hot path
 * ============================================================================================= */

/* Forward-model operators (fwd_model_1d forward_models.py:20-39, fwd_model_2d forward_models.py:57-81) as weight matrices:
 *   1-D: W[i][k]         = scale * b_fwd_1d(z_i - x_k, R) * trapz_w(x)_k          (scale = R / (2 varsigma), :39)
 *   2-D: W[i][a*nx2 + b] = b_fwd_2d(z_i0 - x1_a, z_i1 - x2_b, R, eps) * trapz_w(x1)_a * trapz_w(x2)_b
 * so that LFP = W * CSD is ONE gpcsd_dgemm instead of the reference's Python loop over time x location. */
int gpcsd_fwd_operator_1d(int nz, const double* z, int nx, const double* x, double R, double scale, double* W, long ld,
                          void* stream);
int gpcsd_fwd_operator_2d(int nz, const double* z /* [nz][2] */, int nx1, const double* x1, int nx2, const double* x2, double R,
                          double eps, double* W, long ld, void* stream);

/* Lower Cholesky factor in place (np.linalg.cholesky of sample_prior, gpcsd1d.py:303-304 / gpcsd2d.py:343-350): on return the
 * lower triangle of L holds the factor and the strict upper triangle is zero.  info (device int): 0 = ok, k+1 = the leading
 * minor of order k+1 is not positive definite (numpy raises LinAlgError; the Python layer does the same). */
int gpcsd_cholesky(int n, double* L, long ld, int* info, void* stream);

/* Standard-normal generator for sample_prior (np.random.normal of gpcsd1d.py:308 / gpcsd2d.py:353): Philox4x32-10 counter-based
 * bits + Box-Muller.  Logical element i = row*ncols + col of out[nrows][ld] is the (i & 1)-th normal of counter i >> 1 for the
 * given (seed, stream_id), independent of the launch geometry.  accumulate == 0: out = sd * z;  1: out += sd * z (additive
 * observation noise without materialising it). */
int gpcsd_randn(long nrows, long ncols, long ld, unsigned long long seed, unsigned int stream_id, double sd, int accumulate,
                double* out, void* stream);
/* the generator's raw 4 x 32-bit outputs for counters 0..ncounters-1 (known-answer tests) */
int gpcsd_philox_raw(long ncounters, unsigned long long seed, unsigned int stream_id, unsigned int* out, void* stream);

/* Per-trial evoked-shift objective (auditory_lfp/fit_mean_function.py:311-321), all trials of a batch at once:
 *   gpcsd_shift_residual : Rout[i][j][r] = Y[i][j][r] - mu[0][i][j] - sum_s lerp(mu[s+1][i][:], t_j + tau[r][s])
 *                          (scipy interp1d linear, fill_value="extrapolate"); mu: [nseg+1][nx][nt], tau: [ntrials][nseg]
 *   gpcsd_quad_per_trial : out[r] = sum_ij B[i][j][r]^2 / rD[i][j]  (= sum alpha_r^2 / Dvec with B = alpha / D from
 *                          gpcsd_project_quad); ws: gpcsd_per_trial_ws_doubles(nx, nt, ntrials, ldn, 1)
 *   gpcsd_shift_grad     : out[r][s] = - sum_ij V[i][j][r] * d/dtau lerp(mu[s+1][i][:], t_j + tau[r][s]),  V = K^-1 resid;
 *                          ws: gpcsd_per_trial_ws_doubles(nx, nt, ntrials, ldn, nseg) */
int gpcsd_shift_residual(int nx, int nt, int ntrials, long ldn, const double* Y, int nseg, const double* mu, const double* t,
                         int t_uniform, const double* tau, double* Rout, void* stream);
long gpcsd_per_trial_ws_doubles(int nx, int nt, int ntrials, long ldn, int ncomp);
int gpcsd_quad_per_trial(int nx, int nt, int ntrials, long ldn, const double* B, const double* rD, long ldrd, double* ws,
                         double* out, void* stream);
int gpcsd_shift_grad(int nx, int nt, int ntrials, long ldn, const double* V, int nseg, const double* mu, const double* t,
                     int t_uniform, const double* tau, double* ws, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPCSD_B200_H */
