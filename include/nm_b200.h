/* nm_b200.h -- status-returning C ABI of libnm_b200.so (B200 / sm_100a implementation of the
 * NormalModes hot path: distributed CSR SpMV, ChebIter, Chebyshev-filtered Lanczos, FE assembly).
 * Plain pointers and sizes only.  Every function returns 0 on success, non-zero on error with the
 * message available from nm_last_error_message().  Unless a name ends in _dev, vector arguments
 * are HOST pointers and the call synchronises; _dev variants take DEVICE pointers and are
 * asynchronous on nm_stream().
 *
 * Each group cites the reference interface it replaces (file:line into /root/reference/src).
 * The by-reference Fortran twins the reference actually links are declared in pevsl_f90.h.
 */
#ifndef NM_B200_H
#define NM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*nm_matvec_fn)(double* x, double* y, void* data);

/* ---- errors ---------------------------------------------------------------------------------- */
int nm_last_error(void);
const char* nm_last_error_message(void);
void nm_clear_error(void);

/* ---- runtime: one process per GPU ------------------------------------------------------------- */
int nm_init(int device);                                   /* device < 0: keep the current CUDA device */
int nm_device_info(int* device, int* sm_count, int* rank, int* nranks);
long long nm_launch_count(void);                           /* kernels launched by this library so far */
void* nm_stream(void);                                     /* cudaStream_t all work is queued on */
int nm_sync(void);
/* NCCL communicator replacing mymatvec%comm (mod_matvec.f90:75): rank 0 creates the id, the host
 * broadcasts the 128 bytes (MPI_Bcast / torch.distributed), every rank calls nm_comm_init. */
int nm_comm_unique_id(char* id128);
int nm_comm_init(int rank, int nranks, const char* id128);
int nm_comm_finalize(void);

/* ---- distributed CSR: pevsl_parcsrcreate_f90 / pevsl_parcsrmatvec_f90 (mod_matvec.f90:69,453) ---- */
int nm_parcsr_create(int nrow_glob, int ncol_glob, const int* row_starts, const int* col_starts, const int* ia,
                     const int* ja, const double* a, void** mat_out);
int nm_parcsr_free(void* mat);
int nm_parcsr_matvec(void* mat, const double* x, double* y);
int nm_parcsr_matvec_dev(void* mat, const double* x_dev, double* y_dev);
/* format: 0 CSR, 1 ROW3 (shared column triples), 2 KRON3 (M (x) I3); fmt_bytes: matrix bytes one SpMV streams */
int nm_parcsr_info(void* mat, int* nrow, int* ncol, long long* nnz, int* format, int* nghost, long long* fmt_bytes);

/* Bdiagscaling / Apdiagscaling (mod_matvec.f90:252-342,345-441) on the device, in place:
 * d_i = 1/sqrt(sign*M_ii), M~_ij = (sign*M_ij*d_j)*d_i ; d_host[n_local] receives d (may be NULL). */
int nm_parcsr_jacobi_scale(void* mat, double sign, double* d_host);
int nm_parcsr_get_values(void* mat, double* a_host /* nnz_local */);
/* the halo (ghost-DOF) exchange of one product alone, on device vectors; mode: 0 none, 1 NCCL send/recv, 2 NVLink peer
   window (direct peer stores + arrival flags) */
int nm_parcsr_halo_exchange_dev(void* mat, const double* x_dev);
int nm_parcsr_halo_info(void* mat, int* mode, int* nghost, int* nsend);
/* host-only (no GPU, no communicator): receive side of one rank's halo plan -- sorted ghost column ids (ghost_glob may
   be null) and the number owned by each rank; the device plan of nm_parcsr_create is built from the same routine */
int nm_halo_plan_host(int nranks, int rank, const int* col_starts, long long nnz, const int* ja, int* nghost,
                      int* ghost_glob, int* recv_cnt);

/* ---- ChebIter: pevsl_setup_chebiter_f90 / pevsl_chebiter_f90 (mod_matvec.f90:93,174,480,512) ---- */
int nm_chebiter_create(double lmin, double lmax, int deg, void* mat, void** cheb_out);
int nm_chebiter_free(void* cheb);
int nm_chebiter_solve_host(void* cheb, const double* b, double* x);
int nm_chebiter_solve_dev(void* cheb, const double* b_dev, double* x_dev);
int nm_chebiter_stats(void* cheb, long long* nsolve, long long* nmatvec, int* deg, double* lmin, double* lmax);
/* kind: 0 plain subwarp kernels (fallback), 1-2 retired, 3 warp-sliced ELL slabs (k_slab), 4 the same with
   producer/consumer warps (k_slabws, one launch per step), 5 the whole iteration in one persistent cooperative launch
   (k_slabpers, NM_SLAB_PERS=1: grid barrier or per-chunk flags between steps); on several GPUs kinds 4 and 5 exchange the
   halo inside the kernel through flag-in-data slots;
   bytes: matrix bytes one step streams */
int nm_chebiter_pack_info(void* cheb, int* kind, long long* bytes);
/* diagnostic (NM_SLAB_TRACE=1): per CTA and chunk 8 clock64 stamps of the last k_slab launch, [grid][64][8] */
int nm_chebiter_trace_dump(void* cheb, long long* out, long long cap, int* grid, int* cta_first /* grid+1 */);
/* host-only self-test of the slab packer (no GPU): packs the pattern, walks the blobs as k_slab does and returns
   y = A x in pack order, the pack order and {nchunk, grid, threads, smem_bytes, nstage, max_chunks_per_cta, padded} */
int nm_slab_host_selftest(int n, int ncolb, int R, const int* rp, const int* idx, const double* vals, const double* x,
                          double* y, int* order_out, int* info);

/* ---- operators: the callbacks sparseAV / sparsefsAV / sparseBV / sparseApV (mod_matvec.f90:445-520) ---- */
int nm_op_create_csr(void* mat, void** op_out);                               /* w = M v                        */
int nm_op_create_solid(void* A, const double* d, void** op_out);              /* w = D A D v          (:445-458) */
int nm_op_create_fluidsolid(void* Ad, void* E, void* ET, void* chebAp, const double* d, const double* dp,
                            void** op_out);                                   /* w = D[Ad+E Dp Ap~^-1 Dp ET]D v (:498-520) */
int nm_op_create_callback(int n_local, nm_matvec_fn fn, void* data, void** op_out);
int nm_op_free(void* op);
int nm_op_apply_host(void* op, const double* x, double* y);
int nm_op_apply_dev(void* op, const double* x_dev, double* y_dev);

/* ---- filter polynomial: pevsl_findpol_f90 / pevsl_freepol_f90 (mod_pevsl.f90:115,215); host only ---- */
int nm_findpol_create(const double* xintv4, double thresh_int, double thresh_ext, void** pol_out);
int nm_pol_info(void* pol, int* deg, double* cc, double* dd, double* gam, double* bar, int* type);
int nm_pol_coeffs(void* pol, double* mu /* deg+1 */);
int nm_pol_free(void* pol);
int nm_tridiag_eig_host(int k, const double* d, const double* e, double* w, double* Z, double* lastrow);
/* dense symmetric-definite pencil H c = w G c (m x m, column-major; Cholesky + cyclic Jacobi): the Rayleigh-Ritz
 * refinement of the accepted eigenvectors; host only */
int nm_sym_geneig_host(int m, const double* H, const double* G, double* w, double* C);

/* ---- FE assembly: cg_create_matrix (mod_cg_create_matrix.f90:35-61) ----
 * nm_fem_create  = topology + DOF numbering + CSR patterns on the host (matrixstruct :1269-1455,
 *                  matrixstruct_general :1458-2033; vstat/v2v/P2 edge nodes of mod_geometry.f90:284-689);
 *                  integer-exact with the reference for a given part[] (NULL: nproc = 1).  No GPU needed.
 * nm_fem_assemble_values = element integration + scatter on the device (CGE3D_ISO :980-1266,
 *                  CGFSE3D_ISO :103-977).  Inputs in the reference's file layout but
 *                  0-based: ele/neigh [ntet][4] (neigh -1 = boundary), node [nvert][3],
 *                  vp/vs/rho [ntet][pNp], g0 [ntet][pNp][3] in m/s^2 (JOB >= 2).
 * matrix ids: 0 = A (Ad in the fluid case), 1 = B, 2 = E, 3 = ET, 4 = Ap (CGM%Ap, i.e. before the sign flip
 * of mod_matvec.f90:137).  Columns are 0-based global ids, rows are this rank's rows. */
int nm_fem_create(int ntet, int nvert, const int* ele, const int* neigh, const double* node, int porder,
                  const double* vs, int nproc, const int* part /* [nn] or NULL */, int rank, void** fem_out);
int nm_fem_free(void* fem);
int nm_fem_info(void* fem, int* nn, int* N, int* Np, int* fluidcase, int* n_local_elements);
int nm_fem_matrix_sizes(void* fem, int which, int* present, int* nrow_local, long long* nnz_local);
int nm_fem_matrix_get(void* fem, int which, int* rowdist, int* coldist, int* ia, int* ja, double* val);
int nm_fem_numbering(void* fem, int* vstat, int* vnum, int* pnum, int* vstt, int* pstt, int* order);
int nm_fem_t2n(void* fem, int* t2n);
/* reference-element matrices (host): M [pNp^2], D [3][pNp^2], MF [4][Nfp^2], Fmask [4][Nfp] (src/mod_geometry.f90:2307-2523) */
int nm_refelem_get(int porder, double* M, double* D, double* MF, int* Fmask);
int nm_fem_assemble_values(void* fem, int job, const double* vp, const double* vs, const double* rho,
                           const double* g0);

/* ---- solver context: pevsl_start .. pevsl_copy_result (mod_pevsl.f90:54-130) ---- */
int nm_pevsl_create(void** pevsl_out);
int nm_pevsl_free(void* pevsl);
int nm_pevsl_setprobsizes(void* pevsl, int N_global, int n_local, int nfirst);
int nm_pevsl_set_nfirst(void* pevsl, int nfirst);          /* first global row of this rank (start-vector keying) */
int nm_pevsl_setamv_callback(void* pevsl, nm_matvec_fn fn, void* data);
int nm_pevsl_setbmv_callback(void* pevsl, nm_matvec_fn fn, void* data);
int nm_pevsl_setamv_op(void* pevsl, void* op);
int nm_pevsl_setbmv_op(void* pevsl, void* op);
int nm_pevsl_adopt_op(void* pevsl, void* op);
int nm_pevsl_setbsol_chebiter(void* pevsl, void* cheb);
int nm_pevsl_set_geneig(void* pevsl);
int nm_pevsl_set_seed(void* pevsl, unsigned long long seed);
/* optional, off by default (0 = pEVSL's trace test only): keep iterating until every wanted Ritz pair's Lanczos
 * residual estimate |beta_k s_ki| is below tol (the pairs next to the band edges converge last) */
int nm_pevsl_set_ritz_tol(void* pevsl, double tol);
int nm_pevsl_lanbounds(void* pevsl, int mlan, int lanstep, double tol, double* lmin, double* lmax);
int nm_pevsl_cheblannr(void* pevsl, const double* xintv4, int maxit, double tol, void* pol);
int nm_pevsl_get_nev(void* pevsl, int* nev);
int nm_pevsl_copy_result(void* pevsl, double* vals, double* vecs, int ld, double* res /* may be NULL */);
int nm_pevsl_stats(void* pevsl, int* steps, int* deg, double* t_total, double* t_filter, double* t_reorth,
                   double* t_ritz, long long* n_filter);
/* diagnostic: microseconds of the dense kernels of the Lanczos phase on random data of local length n -- one CGS pass
 * over k basis columns (us[0] c = V^T z incl. the partial reduction, us[1] z -= Z c), us[2] the Ritz product U = V S
 * (k x ns, DMMA), us[3] the Gram block U^T U (ns x ns, DMMA) */
int nm_diag_lanczos_kernels(long long n, int k, int ns, double* us4);
/* one application y = p(A B^-1) z of the polynomial filter (ChebAv), the unit of the headline metric */
int nm_pevsl_filter_host(void* pevsl, void* pol, const double* z, double* y);
int nm_pevsl_filter_dev(void* pevsl, void* pol, const double* z_dev, double* y_dev, double* work3n_dev);
/* the same sum truncated after kmax degree steps (kmax <= 0: the whole polynomial): a fixed-size slice of one
 * application, the unit bench.py times on meshes where a whole application takes minutes */
int nm_pevsl_filter_steps_host(void* pevsl, void* pol, int kmax, const double* z, double* y);
int nm_pevsl_filter_steps_dev(void* pevsl, void* pol, int kmax, const double* z_dev, double* y_dev, double* work3n_dev);

#ifdef __cplusplus
}
#endif
#endif
