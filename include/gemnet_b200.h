/*
 * gemnet_b200 — C ABI of the B200-native (sm_100a) GemNet interaction-block hot path.
 *
 * The reference (TUM-DAML/gemnet_pytorch @ a0164f7) is pure Python and has no FFI; its seams are
 * Python classes.  Each entry point below replaces one reference layer call (cited per function,
 * paths relative to the reference root) and is what a `torch.autograd.Function` (or any other host
 * language through ctypes/cffi) binds.  See INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *    float data is fp32 row-major, index data is int32; sizes are int64_t;
 *  - `stream` is a `cudaStream_t` passed as void*; all work is stream-ordered, nothing allocates,
 *    nothing synchronises: safe under CUDA-graph capture and autograd worker threads;
 *  - return value 0 = ok, non-zero = error (bad argument or CUDA launch error); the message is in
 *    `gemnet_last_error()` (thread-local).  No exceptions cross the ABI;
 *  - `*_bwd` functions ACCUMULATE (+=) into gradient outputs that are documented "(accumulated)"
 *    and overwrite the others; a NULL gradient output pointer skips that gradient.
 */
#ifndef GEMNET_B200_H
#define GEMNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEMNET_B200_ABI_VERSION 1

int gemnet_abi_version(void);
const char* gemnet_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Index preparation (device).  The reference pads every segment to Kmax using `id_reduce`/`Kidx`
 * (efficient.py:173-175, basis_layers.py:153-159); this library uses CSR row pointers instead.
 * ---------------------------------------------------------------------------------------------- */

/* rowptr[r] = first position p with sorted_ids[p] >= r, r in [0, n_rows]; sorted_ids must be
 * non-decreasing (data_container.py:324-332 sorts id3_reduce_ca / id4_reduce_ca).  `flag` (int32[1],
 * zero-initialised by the caller) is set to 1 if the ids are not sorted or out of range. */
int gemnet_csr_from_sorted(const int32_t* sorted_ids, int64_t n, int64_t n_rows, int32_t* rowptr,
                           int32_t* flag, void* stream);

/* Stable counting sort of `ids` (values in [0, n_rows)): perm = positions ordered by (id, position),
 * rowptr as above.  workspace: int32[n_rows + 1 + n] (cursor + unordered placement).  Used for the edge->atom CSR (id_a) and the
 * transposed triplet CSR (id3_expand_ba).  An id outside [0, n_rows) is skipped and sets `flag` (int32[1],
 * zero-initialised by the caller, may be NULL) to 1: a malformed index list is reported, never written out of bounds. */
int gemnet_csr_by_key(const int32_t* ids, int64_t n, int64_t n_rows, int32_t* rowptr, int32_t* perm,
                      int32_t* workspace, int32_t* flag, void* stream);

/* Range check of the remaining gather lists (id3_expand_ba, id4_expand_*, id4_int_*, ... : the reference indexes
 * tensors with them and torch raises on an invalid index; here the kernels would read out of bounds): `flag` (int32[1],
 * zero-initialised by the caller) is set to 1 if any id of list l lies outside [0, bounds_host[l]).  The three arrays
 * are HOST arrays of n_lists <= GEMNET_RANGE_CHECK_MAX_LISTS entries (device list pointers, sizes, bounds); one launch. */
#define GEMNET_RANGE_CHECK_MAX_LISTS 12
int gemnet_index_range_check(const int32_t* const* lists_host, const int64_t* sizes_host, const int64_t* bounds_host,
                             int32_t n_lists, int32_t* flag, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H1  GemNet.calculate_interatomic_vectors (gemnet/model/gemnet.py:262-286)
 *     D[e] = |R[id_a[e]] - R[id_c[e]]|,  V[e,:] = (R[id_a[e]] - R[id_c[e]]) / D[e]
 * bwd: gR (A,3) (accumulated, atomics) from gD (E) and gV (E,3) (either may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int gemnet_edge_geometry_fwd(const float* R, const int32_t* id_c, const int32_t* id_a, int64_t n_edges,
                             float* D, float* V, void* stream);
int gemnet_edge_geometry_bwd(const float* D, const float* V, const int32_t* id_c, const int32_t* id_a,
                             int64_t n_edges, const float* gD, const float* gV, float* gR, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H3  BesselBasisLayer.forward (gemnet/model/layers/basis_layers.py:45-49) + Envelope (envelope.py:22-29)
 *     rbf[e,n] = env(d/c) * sqrt(2/c) * sin(freq[n] * d/c) / d
 * bwd: gD (E) overwritten; gfreq (num_radial) accumulated (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int gemnet_bessel_rbf_fwd(const float* D, const float* freq, int64_t n_edges, int32_t num_radial,
                          float cutoff, int32_t envelope_p, float* rbf, void* stream);
int gemnet_bessel_rbf_bwd(const float* D, const float* freq, int64_t n_edges, int32_t num_radial,
                          float cutoff, int32_t envelope_p, const float* g_rbf, float* gD, float* gfreq,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * H4/H5 radial part: SphericalBasisLayer / TensorBasisLayer `rbf_env`
 *     (basis_layers.py:121-128, 241-250; formulas basis_utils.py:47-80)
 *     rad[e,l,n] = env(d/c) * c^-1.5 * norm[l,n] * j_l(zeros[l,n] * d/c)
 * `zeros` / `norm` are (num_spherical, num_radial) float64 DEVICE tables built by the host
 * (float32-rounded zeros as the reference bakes into its formulas).  Evaluated in fp64, stored fp32.
 * Layout (E, num_spherical, num_radial) — the reference's (num_spherical, E, num_radial) transposed.
 * ---------------------------------------------------------------------------------------------- */
int gemnet_sph_bessel_radial_fwd(const float* D, int64_t n_edges, int32_t num_spherical, int32_t num_radial,
                                 const double* zeros, const double* norm, float cutoff, int32_t envelope_p,
                                 float* rad, void* stream);
int gemnet_sph_bessel_radial_bwd(const float* D, int64_t n_edges, int32_t num_spherical, int32_t num_radial,
                                 const double* zeros, const double* norm, float cutoff, int32_t envelope_p,
                                 const float* g_rad, float* gD, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H2+H4 angular part, triplets: GemNet.calculate_angles3 (gemnet.py:421-451) fused with the Y_l0
 * evaluation of SphericalBasisLayer (basis_layers.py:130-131; basis_utils.py:131-138,214-215).
 *     u = R[id_c[red]] - R[id_a[red]], v = R[id_c[exp]] - R[id_a[red]]
 *     x = u.v, y = max(|u x v|, 1e-9), cos = x / sqrt(x^2+y^2)       (== cos(atan2(y, x)), gemnet.py:303-311)
 *     sph[t,l] = pref[l] * P_l(cos),  pref[l] = sqrt((2l+1)/4pi) (float32[num_spherical], device)
 * `angle` (T) optional output (atan2(y,x)) for parity checks of row H2.  No padding: row t of `sph`
 * is triplet t in the caller's (CSR-sorted) order.
 * bwd: gR (A,3) accumulated (atomics).
 * ---------------------------------------------------------------------------------------------- */
int gemnet_triplet_sph_fwd(const float* R, const int32_t* id_c, const int32_t* id_a, const int32_t* id3_reduce_ca,
                           const int32_t* id3_expand_ba, int64_t n_triplets, int32_t num_spherical,
                           const float* pref, float* sph, float* angle, void* stream);
int gemnet_triplet_sph_bwd(const float* R, const int32_t* id_c, const int32_t* id_a, const int32_t* id3_reduce_ca,
                           const int32_t* id3_expand_ba, int64_t n_triplets, int32_t num_spherical,
                           const float* pref, const float* g_sph, float* gR, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H2+H4 (cbf4) / H2+H5 (sbf4), quadruplets: GemNet.calculate_angles (gemnet.py:335-418) fused with
 *   - SphericalBasisLayer(efficient=False) angular part on intermediate triplets a<-b->d  (basis_layers.py:130-144)
 *   - TensorBasisLayer real Y_lm(theta=Phi_cab, phi=Theta_cabd)                            (basis_layers.py:269-270)
 * Stage 1 (per intermediate triplet j of the a-b<-d side):
 *     ba = R[a]-R[b], bd = R[d]-R[b]  (a=id4_int_a[id4_expand_intm_ab[j]], b=id4_int_b[...], d=id_c[id4_expand_intm_db[j]])
 *     sph_abd[j,l] = pref0[l] P_l(cos(angle(ba,bd)));  bd_proj[j,:] = bd - (bd.ba/ba.ba) ba
 * Stage 2 (per intermediate triplet j of the c->a-b side):
 *     ac = R[c]-R[a], ab = R[b]-R[a]  (c=id_c[id4_reduce_intm_ca[j]], a=id_a[...], b=id4_int_b[id4_reduce_intm_ab[j]])
 *     cs_cab[j,:] = (cos, sin) of angle(ab,ac);       ac_proj[j,:] = ac - (ac.ab/ab.ab) ab
 * Stage 3 (per quadruplet q): theta from cs_cab[id4_reduce_cab[q]], phi = angle(ac_proj[id4_reduce_cab[q]], bd_proj[id4_expand_abd[q]])
 *     ylm[q, :] in the reference order: l ascending, within l: m = 0, +1..+l, -l..-1 (basis_utils.py:227-243)
 * `pref_lm` float32[num_spherical^2] (device): sqrt(2)^(m!=0) (-1)^m sqrt((2l+1)/4pi (l-|m|)!/(l+|m|)!) in that order.
 * ---------------------------------------------------------------------------------------------- */
int gemnet_quad_intm_fwd(const float* R, const int32_t* atom_center, const int32_t* atom_first,
                         const int32_t* atom_second, int64_t n_intm, int32_t num_spherical, const float* pref0,
                         float* sph /* (M,S) or NULL */, float* cossin /* (M,2) or NULL */,
                         float* proj /* (M,3) rejection of (second-center) on (first-center) */, void* stream);
int gemnet_quad_intm_bwd(const float* R, const int32_t* atom_center, const int32_t* atom_first,
                         const int32_t* atom_second, int64_t n_intm, int32_t num_spherical, const float* pref0,
                         const float* g_sph, const float* g_cossin, const float* g_proj, float* gR, void* stream);
int gemnet_quad_ylm_fwd(const float* cossin_cab, const float* ac_proj, const float* bd_proj,
                        const int32_t* id4_reduce_cab, const int32_t* id4_expand_abd, int64_t n_quads,
                        int32_t num_spherical, const float* pref_lm, float* ylm, float* theta_phi /* (Q,2) or NULL */,
                        void* stream);
int gemnet_quad_ylm_bwd(const float* cossin_cab, const float* ac_proj, const float* bd_proj,
                        const int32_t* id4_reduce_cab, const int32_t* id4_expand_abd, int64_t n_quads,
                        int32_t num_spherical, const float* pref_lm, const float* g_ylm,
                        float* g_cossin_cab /* (M,2) accumulated */, float* g_ac_proj /* (M,3) accumulated */,
                        float* g_bd_proj /* (M,3) accumulated */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H6  EfficientInteractionDownProjection.forward (gemnet/model/layers/efficient.py:41-57)
 *     rbf_W1[e,i,s] = sum_n rad[e, l_of_s[s], n] * Wd[s,n,i]
 * `l_of_s` int32[S]: identity for triplets (S = num_spherical); for quadruplets S = num_spherical^2 and
 * l_of_s implements torch.repeat_interleave(rbf_env, 2l+1) (basis_layers.py:251-253) without materialising it.
 * bwd: g_rad (E,L,R) overwritten; g_Wd (S,R,I) accumulated (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int gemnet_downproj_fwd(const float* rad, const float* Wd, const int32_t* l_of_s, int64_t n_edges, int32_t L,
                        int32_t num_radial, int32_t S, int32_t I, float* rbf_W1, void* stream);
int gemnet_downproj_bwd(const float* rad, const float* Wd, const int32_t* l_of_s, int64_t n_edges, int32_t L,
                        int32_t num_radial, int32_t S, int32_t I, const float* g_rbf_W1, float* g_rad,
                        float* g_Wd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H7  EfficientInteractionBilinear.forward (gemnet/model/layers/efficient.py:159-189) fused with the
 *     source-edge gather of TripletInteraction / QuadrupletInteraction (interaction_block.py:678 / 562)
 *     and the ScalingFactor multiply that follows it (interaction_block.py:682 / 555; scaling.py:170-174):
 *
 *     out[e,o] = scale * sum_{t in [rowptr[e], rowptr[e+1])} sum_{c,i,s}
 *                        x[expand[t], c] * sph[t,s] * rbf_W1[e,i,s] * W[c,i,o]
 *
 * A ragged (CSR) segmented GEMM: no (E,Kmax,.) padding, no gathered (T,C) tensor in HBM.
 * Supported shapes (S,C,I,O): (7,64,16,64) triplets and (49,32,32,32) quadruplets of the published
 * configuration, plus the generic fallback kernel for any S<=64, C in {16,32,64,128}, I<=64, O<=128.
 * `algo`: 0 = fp32 CUDA-core kernels; GEMNET_SEGBIL_TC (1) = tcgen05 kernels where built — the triplet shape: the
 * forward's (E x 1024) @ (1024 x 64) stage and the backward's dB = g W'^T stage (the quadruplet shape and `algo` 1 fall
 * back to the fp32 kernels in the backward of other shapes); | GEMNET_SEGBIL_PREPACKED (2): `W` points to the image
 * written by gemnet_segbilinear_pack (forward image for _fwd, backward image for _bwd) instead of the raw weight, so
 * that the per-call pack kernel disappears (weights are packed once per version by the caller).
 * bwd: g_x (n_x, C) accumulated (atomics); g_sph (U,S), g_rbf_W1 (E,I,S) overwritten;
 *      g_W (C,I,O) accumulated (may be NULL: inference / force evaluation does not need it).
 * workspace: gemnet_segbilinear_workspace_bytes() bytes (may return 0).
 * ---------------------------------------------------------------------------------------------- */
#define GEMNET_SEGBIL_TC 1
#define GEMNET_SEGBIL_PREPACKED 2
/* _bwd only, fp32 split pipeline (the quadruplet shape): the workspace handed in is the one the forward call used
 * (allocated with the BACKWARD size), whose head still holds the segment sums A (E,S,C): the backward skips
 * recomputing them (166 us of a 700 us quadruplet backward at batch 64) */
#define GEMNET_SEGBIL_A_SAVED 4
int64_t gemnet_segbilinear_workspace_bytes(int64_t n_edges, int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo,
                                           int32_t backward);
/* packed bf16 hi/lo weight images of the tcgen05 kernels (0 bytes = no tensor-core kernel for the shape) */
int64_t gemnet_segbilinear_pack_bytes(int32_t S, int32_t C, int32_t I, int32_t O, int32_t backward);
int gemnet_segbilinear_pack(const float* W, int32_t S, int32_t C, int32_t I, int32_t O, int32_t backward, void* packed,
                            void* stream);
/* phase 1 on its own: A (E, S, C) = sum over the segment of sph[t,s] * x[expand[t],c] (training: dL/dW of the layer is
 * two library contractions on top of it; no (U, S, C) intermediate) */
int gemnet_segbilinear_segsum(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                              int64_t n_edges, int32_t S, int32_t C, int32_t I, int32_t O, float* A_out, void* stream);
int gemnet_segbilinear_fwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                           const float* rbf_W1, const float* W, float scale, int64_t n_edges, int64_t n_units,
                           int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo, float* out, void* workspace,
                           void* stream);
int gemnet_segbilinear_bwd(const float* x, const int32_t* expand, const int32_t* rowptr, const float* sph,
                           const float* rbf_W1, const float* W, float scale, int64_t n_edges, int64_t n_units,
                           int32_t S, int32_t C, int32_t I, int32_t O, int32_t algo, const float* g_out,
                           float* g_x, float* g_sph, float* g_rbf_W1, float* g_W, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H10/H11  edge->atom aggregation of AtomUpdateBlock / OutputBlock
 *     (gemnet/model/layers/atom_update_block.py:64-68 and :166-173; torch_scatter.scatter reduce="add")
 *     out[a,f] = scale * sum_{e in in(a)} m[e,f] * r[e,f]          in(a) = perm[rowptr[a] .. rowptr[a+1])
 * deterministic (no atomics): edges of an atom are summed in ascending edge id.
 * `had` (E,F) optional: had[e,f] = had_scale * m[e,f] * r[e,f]   (OutputBlock force branch, :181-182)
 * ---------------------------------------------------------------------------------------------- */
int gemnet_edge_to_atom_fwd(const float* m, const float* r, const int32_t* perm, const int32_t* rowptr,
                            int64_t n_atoms, int64_t n_edges, int32_t F, float scale, float* out, float had_scale,
                            float* had, void* stream);
int gemnet_edge_to_atom_bwd(const float* m, const float* r, const int32_t* id_a, int64_t n_edges, int32_t F,
                            float scale, const float* g_out, float had_scale, const float* g_had, float* g_m,
                            float* g_r, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H12  read-out of GemNet.forward (gemnet/model/gemnet.py:578-597)
 *   molecule energy: E[mol,t] = sum|mean over atoms of the molecule (atoms of a molecule are contiguous,
 *                    batch_seg non-decreasing, data_container.py:235) — mol_rowptr from gemnet_csr_from_sorted
 *   direct forces  : F[a,t,:] = sum_{e in in(a)} F_ca[e,t] * V[e,:]
 * ---------------------------------------------------------------------------------------------- */
int gemnet_molecule_sum_fwd(const float* E_atom, const int32_t* mol_rowptr, int64_t n_mol, int32_t n_targets,
                            int32_t mean, float* E_mol, void* stream);
int gemnet_molecule_sum_bwd(const float* g_E_mol, const int32_t* batch_seg, const int32_t* mol_rowptr,
                            int64_t n_atoms, int32_t n_targets, int32_t mean, float* g_E_atom, void* stream);
int gemnet_force_readout_fwd(const float* F_ca, const float* V, const int32_t* perm, const int32_t* rowptr,
                             int64_t n_atoms, int32_t n_targets, float* F_atom, void* stream);
int gemnet_force_readout_bwd(const float* F_ca, const float* V, const int32_t* id_a, int64_t n_edges,
                             int32_t n_targets, const float* g_F_atom, float* g_F_ca, float* g_V, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1 (first step): fused pointwise kernels around the dense layers and gathers with a deterministic backward.
 *   scaled_silu  mode 0: out = silu(z)/0.6                       ScaledSiLU (base_layers.py:51-58)
 *                mode 1: out = (x + silu(z)/0.6) * s             tail of ResidualLayer (base_layers.py:86-89)
 *   scaled_silu_bwd      gz = g * s * d/dz[silu(z)/0.6]          (the x-branch gradient of mode 1 is g*s: add_scale)
 *   add_scale            out = (a [+ b [+ c]]) * s               skip connections (interaction_block.py:202-233)
 *   mul_scale            out = a * b * s                         Hadamard + ScalingFactor (interaction_block.py:661-664)
 *   gather_rows          out[j,:] = x[idx[j],:]                  h[id_a], h[id_c], x[id_swap] (embedding_block.py:70-71)
 *   segment_sum_rows     out[r,:] = sum_k src[perm[k],:] over the CSR row r — the gather's backward, no atomics
 * n = number of floats (any; the first n - n%4 are processed as 128-bit vectors); pointers 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int gemnet_scaled_silu_fwd(const float* z, const float* x, float s, int64_t n, int32_t mode, float* out, void* stream);
int gemnet_scaled_silu_bwd(const float* z, const float* g, float s, int64_t n, float* gz, void* stream);
int gemnet_add_scale(const float* a, const float* b, const float* c, float s, int64_t n, float* out, void* stream);
/* out = sa*a [+ sb*b [+ sc*c]] (b, c may be NULL): gradient accumulation of tensors with several consumers, with the
 * 1/sqrt(2), 1/sqrt(3) factors of the skip connections (interaction_block.py:202-233) folded in. */
int gemnet_lincomb3(const float* a, float sa, const float* b, float sb, const float* c, float sc, int64_t n, float* out,
                    void* stream);
int gemnet_mul_scale_fwd(const float* a, const float* b, float s, int64_t n, float* out, void* stream);
int gemnet_mul_scale_bwd(const float* a, const float* b, const float* g, float s, int64_t n, float* ga, float* gb,
                         void* stream);
int gemnet_gather_rows(const float* x, const int32_t* idx, int64_t n, int32_t F, float* out, void* stream);
int gemnet_segment_sum_rows(const float* src, const int32_t* perm, const int32_t* rowptr, int64_t n_rows, int32_t F,
                            float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1: fused dense layer on tcgen05 (Dense / ResidualLayer, gemnet/model/layers/base_layers.py:5-89):
 *     y    = s_acc * f(A @ B^T) + s_res * res,   A = x or x .* xmul,   f = identity (act=0) | silu(.)/0.6 (act=1)
 *     dact = s_acc * f'(A @ B^T)                 (optional output: the factor the backward multiplies the gradient by)
 * fp32 in / fp32 out; both operands are split exactly into three bf16 terms and the six significant products are
 * accumulated in fp32 in TMEM (main and correction sums kept apart, the accumulator's toward-zero rounding
 * compensated), so the result is fp32-grade.  B is given pre-packed: gemnet_dense_pack_weight converts an (N x K) operand
 * with arbitrary element strides (so W, W^T and column slices of W all pack without copies) into TWO images, one after
 * the other in `packed` (gemnet_dense_pack_floats(N, K) floats in total): the resident shared-memory image of
 * gemnet_dense_tc (K-major SWIZZLE_128B chunks of 64 k) and the tensor-memory image of gemnet_chain_tc (u32
 * [split][K / 8 column groups][128 features][4 columns], column c = bf16 pair k = 2c, 2c + 1: that kernel's MMA reads
 * the weights as its A operand from TMEM).  N, K multiples of 4 in [4, 128]; x rows have stride ldx.
 * flags: GEMNET_DENSE_WEIGHTS_SETTLED = the packed image was NOT produced by the kernel launched immediately before
 * this call on `stream` (weights are packed once and reused): the kernel is then launched as a programmatic
 * dependent of its predecessor and fetches the image while the predecessor is still running.
 * ---------------------------------------------------------------------------------------------- */
#define GEMNET_DENSE_WEIGHTS_SETTLED 1
int64_t gemnet_dense_pack_floats(int32_t N, int32_t K);
int gemnet_dense_pack_weight(const float* W, int64_t stride_n, int64_t stride_k, int32_t N, int32_t K, float* packed,
                             void* stream);
/* the same for many weights in ONE launch: `table_device` = n_entries descriptors in DEVICE memory (training re-packs
 * every image after every optimiser step; Np / Kp = N / K rounded up to 64 or 128 as gemnet_dense_pack_weight does) */
typedef struct {
  const float* src;
  float* dst;
  int64_t stride_n, stride_k;
  int32_t N, K, Np, Kp;
} gemnet_pack_desc;
int gemnet_dense_pack_many(const gemnet_pack_desc* table_device, int32_t n_entries, void* stream);
int gemnet_dense_tc(const float* x, const float* xmul, int64_t M, int32_t K, int32_t ldx, const float* packed,
                    int32_t N, int32_t act, float s_acc, const float* res, float s_res, float* y, float* dact,
                    int32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1 (chain fusion): a whole stack of dense layers with the pointwise work between them in ONE kernel —
 * ResidualLayer (base_layers.py:61-89), the dense_ba -> Hadamard -> down_projection prologue and the
 * up_projection / skip / residual tail of TripletInteraction / QuadrupletInteraction / InteractionBlock
 * (interaction_block.py:186-233, 385-421, 517-566, 653-696), the atom MLPs of AtomUpdateBlock / OutputBlock
 * (atom_update_block.py:55-72, 157-193), EdgeEmbedding without the concat (embedding_block.py:70-74) — and the
 * input-gradient of each of them (forces are -dE/dR, gemnet.py:598-613).
 *
 * A 64-row tile (16 rows for atom-level chains) of the activation stays on chip for the whole chain: the operand
 * lives in shared memory as an exact three-way bf16 split (gemnet_dense_tc's numerics), the weights, the accumulators
 * and a skip-connection value ("carry") in tensor memory; only what the backward pass needs (activation derivatives)
 * and the chain's outputs go to HBM.  The two 32-row halves of a tile are software-pipelined: epilogue of one half
 * under the GEMM of the other (csrc/chain_tc.cu).  The chain is described by a small program of up to GEMNET_CHAIN_MAX_STEPS steps; every step is
 *
 *   1. (optional) operand fill   v[r,:] = sum_t scale_t * src_t[idx_t ? idx_t[r] : r, :] (* mul_t[r,:])   t < n_terms
 *                                (kind 1: v[r,c] = scale * src[r] * mul[c], an outer product: the N=1 output layer's
 *                                backward); carry = v if load_set_carry; A = split3(v * amul) (K columns)
 *   2. GEMM                      acc = A @ B^T, B = packed (N x K) image of gemnet_dense_pack_weight
 *   3. epilogue (per element)    z = acc + pre0[pre0_idx ? pre0_idx[r] : r, c] + pre1[...]
 *                                v = act ? silu(z)/0.6 : z ;  d = act ? d/dz : 1
 *                                mul_early:  aux = v*s_acc (if aux) ; v *= mul_early ; d *= mul_early
 *                                v = v*s_acc + carry*s_carry + post0*s_post0 + post1*s_post1 ;  d *= s_acc
 *                                mul_late:   v *= mul_late
 *                                dact = d ; out[out_idx ? out_idx[r] : r] = v * s_store (* smul) ; carry = v * s_set_carry
 *                                next operand: A = split3(v * s_a * next_amul)   (GEMNET_CHAIN_TO_A)
 *                                           or A = split3(carry * s_a * next_amul) (GEMNET_CHAIN_A_FROM_CARRY)
 *                                dot:        dot_out[r] = sum_c v[r,c] * dot_w[c] (+ dot_in[r])   (N = 1 output layer)
 * All tensors are fp32 row-major with row stride = their logical width (K for operand-side tensors, N for
 * epilogue-side tensors); K, N multiples of 8 in [8, 128]; pointers 16-byte aligned.  The first step of a program must
 * have n_terms > 0; a later step with n_terms == 0 uses the operand its predecessor left (TO_A / A_FROM_CARRY, or
 * the unchanged operand).  flags of gemnet_chain_tc: GEMNET_DENSE_WEIGHTS_SETTLED as for gemnet_dense_tc.
 * ---------------------------------------------------------------------------------------------- */
#define GEMNET_CHAIN_MAX_STEPS 12
#define GEMNET_CHAIN_MAX_TERMS 3
#define GEMNET_CHAIN_ACT 1          /* ScaledSiLU */
#define GEMNET_CHAIN_CARRY_ADD 2    /* v += carry * s_carry */
#define GEMNET_CHAIN_SET_CARRY 4    /* carry = v * s_set_carry */
#define GEMNET_CHAIN_TO_A 8         /* next operand from v */
#define GEMNET_CHAIN_A_FROM_CARRY 16 /* next operand from carry */

typedef struct {
  const float* src;
  const int32_t* idx;
  const float* mul;
  float scale;
  int32_t kind; /* 0 rows, 1 outer product */
} gemnet_chain_term;

typedef struct {
  int32_t n_terms;
  int32_t load_set_carry;
  gemnet_chain_term term[GEMNET_CHAIN_MAX_TERMS];
  const float* amul;
  const void* Wp;
  int32_t K, N;
  int32_t flags;
  int32_t reserved;
  const float* pre0;
  const int32_t* pre0_idx;
  const float* pre1;
  const int32_t* pre1_idx;
  const float* mul_early;
  const float* mul_late;
  const float* post0;
  const float* post1;
  float s_acc, s_carry, s_post0, s_post1, s_store, s_set_carry, s_a, s_pad;
  float* aux;
  float* dact;
  float* out;
  const float* smul;
  const int32_t* out_idx;
  const float* next_amul;
  const float* dot_w;
  float* dot_out;
  const float* dot_in;
  float* save_x;  /* optional (M, K): this step's GEMM operand as written (training: the weight gradient's factor) */
} gemnet_chain_step;

typedef struct {
  int64_t M;
  int32_t n_steps;
  int32_t reserved;
  gemnet_chain_step step[GEMNET_CHAIN_MAX_STEPS];
} gemnet_chain_program;

int gemnet_chain_tc(const gemnet_chain_program* program_host, int32_t flags, void* stream);
/* rows per tile of the following launches: 0 (default) = automatic — 16-row tiles while ceil(M / 16) <= the SM count
 * (atom-level chains: one 16-row tile per SM), else 64 (two pipelined 32-row halves); 16 / 64 force one variant
 * (tests, measurements).  Results do not depend on it beyond fp32 summation order of the dot epilogue. */
int gemnet_chain_set_tile_rows(int32_t rows);
/* diagnostics: when non-NULL, thread 0 of CTA 0 of every following gemnet_chain_tc launch writes clock64() stamps of its
 * phase boundaries (<= 250 int64) to this device buffer; NULL (default) switches it off. */
void gemnet_chain_set_profile_buffer(void* device_int64_x256);

/* ------------------------------------------------------------------------------------------------
 * N1: every projection of the radial basis in one launch.  GemNet maps rbf (E, num_radial) through a two-level
 * bias-free linear fan: level 1 (shared) mlp_rbf3 / mlp_rbf_h / mlp_rbf_out [/ mlp_rbf4] (gemnet.py:206-247, 16 wide),
 * level 2 (one per consumer) TripletInteraction.mlp_rbf / QuadrupletInteraction.mlp_rbf (interaction_block.py:458-463,
 * 601-606), AtomUpdateBlock.dense_rbf / OutputBlock.dense_rbf (atom_update_block.py:36-41,128):
 *     out[j] (E, H) = (rbf W1[group[j]]^T) W2[j]^T          j < n_out,  W1[g] (16, num_radial),  W2[j] (H, 16)
 * `out` is one (n_out, E, H) buffer.  W1_host / W2_host / group_host are HOST arrays (of device pointers / ints); the
 * backward takes the output gradients as a HOST array of n_out device pointers (NULL = no gradient for that output):
 *     g_rbf (E, num_radial) = sum_j (g_out[j] W2[j]) W1[group[j]]     deterministic (fixed order over j)
 * workspace: n_out * E * 16 floats.
 * ---------------------------------------------------------------------------------------------- */
#define GEMNET_RBF_FAN_MAX_OUT 24
#define GEMNET_RBF_FAN_MAX_GROUPS 4
int gemnet_rbf_fan_fwd(const float* rbf, int64_t n_edges, int32_t num_radial, const float* const* W1_host,
                       int32_t n_groups, const float* const* W2_host, const int32_t* group_host, int32_t n_out, int32_t H,
                       float* out, void* stream);
int gemnet_rbf_fan_bwd(int64_t n_edges, int32_t num_radial, const float* const* W1_host, int32_t n_groups,
                       const float* const* W2_host, const int32_t* group_host, int32_t n_out, int32_t H,
                       const float* const* g_out_host, float* workspace, float* g_rbf, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H0  host index builder: DataContainer.__getitem__ / get_triplets / get_quadruplets
 *     (gemnet/training/data_container.py:156-408, 411-489).  HOST function, native C++.
 * Input: molecules concatenated, R_host float64 (n_atoms,3), n_per_mol int32[n_mol].
 * Output through an opaque handle; arrays are int32.  Edge order, id_swap, id_undir, id4_int_*, and the
 * intermediate-triplet arrays are identical to the reference; the reduce-sorted arrays (id3_*, id4_*_ca/db/cab/abd)
 * are emitted grouped by the reduce edge with neighbours in ascending (intermediate) position — the reference's
 * order within a segment is unspecified (np.argsort default kind is unstable, data_container.py:326,371).
 * ---------------------------------------------------------------------------------------------- */
typedef struct gemnet_graph gemnet_graph;
int gemnet_graph_build_host(const double* R_host, const int32_t* n_per_mol_host, int64_t n_mol, double cutoff,
                            double int_cutoff, int32_t triplets_only, gemnet_graph** out);
/* number of elements of array `name` ("id_c", "id3_expand_ba", "Kidx4", ... reference key names), -1 if unknown */
int64_t gemnet_graph_size(const gemnet_graph* g, const char* name);
/* copies array `name` into dst_host (int32, gemnet_graph_size elements) */
int gemnet_graph_copy_host(const gemnet_graph* g, const char* name, int32_t* dst_host);
/* the same, widened to the int64 the reference's tensors use (data_container.py:505-512); multi-threaded */
int gemnet_graph_copy_host_i64(const gemnet_graph* g, const char* name, int64_t* dst_host);
void gemnet_graph_free(gemnet_graph* g);

/* ------------------------------------------------------------------------------------------------
 * N2  DEVICE index builder: the same arrays as gemnet_graph_build_host (DataContainer.__getitem__ /
 *     get_triplets / get_quadruplets, gemnet/training/data_container.py:156-408, 411-489), element for element,
 *     produced on the GPU from the coordinates: count -> exclusive scan -> emission at the final position; no
 *     atomics, no sort, deterministic.  Nothing allocates or synchronises: the caller owns every array below and
 *     passes CAPACITIES; the real sizes appear in the device header `hdr`.  Two ways to use it:
 *       dynamic shapes   count -> read hdr -> allocate the edge-level arrays -> edges -> (quadruplet models) read
 *                        hdr[QUADS] -> allocate -> quads
 *       static shapes    (CUDA-graph serving of triplet models) count, edges, pad back to back on capacity-sized
 *                        arrays; hdr[OVERFLOW] != 0 afterwards means a capacity was too small and the index arrays
 *                        were left untouched (gemnet_graph_count_host gives the host the sizes beforehand).
 *     All index arrays are int32.  `hdr` is device int32[GEMNET_GRAPH_HDR_WORDS]:
 *       [0] n_atoms and [1] n_molecules are WRITTEN BY THE CALLER (device memory, so that a captured graph serves
 *       batches of any size); gemnet_graph_device_count writes [2] upper-triangle pairs, [3] n_edges, [4] n_triplets,
 *       [5] interaction edges, [6] intermediate triplets (each of the two lists), [8] overflow bits
 *       (1 atoms/molecules inconsistent, 2 edges, 4 triplets, 8 interaction edges, 16 intermediate, 32 quadruplets,
 *       64 no room for the pad molecule),
 *       [9] bits of max |float(R)|; gemnet_graph_device_edges writes [7] n_quadruplets.
 * ---------------------------------------------------------------------------------------------- */
#define GEMNET_GRAPH_HDR_WORDS 16
typedef struct gemnet_device_graph {
  /* inputs */
  const double* R;           /* (cap_atoms, 3) float64; rows >= n_atoms are not read */
  const int32_t* n_per_mol;  /* (cap_molecules) atoms per molecule; entries >= n_molecules are not read */
  int32_t* hdr;              /* int32[GEMNET_GRAPH_HDR_WORDS], see above */
  double cutoff, int_cutoff;
  int32_t triplets_only;
  /* capacities, in elements; an array that a stage does not touch may be NULL with capacity 0 */
  int64_t cap_atoms, cap_molecules, cap_edges, cap_triplets, cap_int_edges, cap_intm, cap_quads;
  /* workspace */
  int32_t* mol_off;       /* cap_molecules + 1 */
  int32_t* atom_cnt;      /* 3 * cap_atoms */
  int32_t* upper_off;     /* cap_atoms + 1 */
  int32_t* in_ptr;        /* cap_atoms + 1: CSR of the edges by TARGET atom (row x of the reference's edge_ids) */
  int32_t* int_off;       /* cap_atoms + 1 (quadruplet models) */
  int32_t* incoming;      /* cap_edges: edge ids by target atom, ascending source atom */
  int32_t* incoming_pos;  /* cap_edges: position of an edge in its target's list */
  int32_t* intm_off_a;    /* cap_int_edges + 1 */
  int32_t* intm_off_b;    /* cap_int_edges + 1 */
  int32_t* quad_ptr;      /* cap_edges + 1: CSR of the quadruplets by reduce edge (= rowptr of id4_reduce_ca) */
  /* outputs (reference key names, data_container.py:186-228) */
  float* R32;             /* (cap_atoms, 3) float32 copy of R (the model input), may be NULL */
  int32_t* batch_seg;     /* cap_atoms */
  int32_t *id_c, *id_a, *id_undir, *id_swap;            /* cap_edges */
  int32_t *id3_expand_ba, *id3_reduce_ca, *Kidx3;       /* cap_triplets (Kidx3 may be NULL) */
  int32_t* rowptr3;                                     /* cap_edges + 1: CSR of the triplets by reduce edge */
  int32_t *id4_int_a, *id4_int_b;                       /* cap_int_edges */
  int32_t *id4_reduce_intm_ca, *id4_reduce_intm_ab, *id4_expand_intm_db, *id4_expand_intm_ab; /* cap_intm */
  int32_t *id4_reduce_ca, *id4_expand_db, *id4_reduce_cab, *id4_expand_abd, *Kidx4;           /* cap_quads */
} gemnet_device_graph;

/* stage 1: molecule offsets, batch_seg, R32, per-atom neighbour counts, their scans, the sizes in hdr */
int gemnet_graph_device_count(const gemnet_device_graph* g, void* stream);
/* stage 2: edges, incoming lists, triplets (+ rowptr3); quadruplet models: interaction edges, intermediate
 * triplets, quadruplet counts per reduce edge (quad_ptr) and hdr[7] */
int gemnet_graph_device_edges(const gemnet_device_graph* g, void* stream);
/* stage 3 (quadruplet models): the five quadruplet arrays */
int gemnet_graph_device_quads(const gemnet_device_graph* g, void* stream);
/* static-shape serving of triplet models: fills everything behind the real entries with the inert pad pattern of
 * gemnet_pad_batch_host.  Z (cap_atoms, in place) and N_out (cap_molecules + 1) complete the model's input dict. */
int gemnet_graph_device_pad(const gemnet_device_graph* g, int32_t* Z, int32_t* N_out, void* stream);
/* HOST: the edge and triplet counts of a batch (one pass over the per-molecule distance tests, the cheapest part of
 * H0), for choosing the capacity bucket before a static-shape launch */
int gemnet_graph_count_host(const double* R_host, const int32_t* n_per_mol_host, int64_t n_mol, double cutoff,
                            int64_t* n_edges, int64_t* n_triplets);

/* ------------------------------------------------------------------------------------------------
 * N4  static-shape serving: HOST function.  Writes a reference-format triplet batch (int64 index lists as
 * DataContainer / convert_to_tensor produce them, data_container.py:491-518; float32 R) into capacity-sized int32 /
 * float32 staging arrays and appends inert pad entries (a pad molecule with index cap_molecules whose atoms, edges and
 * triplets only reference each other), so that one captured CUDA graph serves batches of any size up to the capacity.
 * Requires n_atoms + 3 <= cap_atoms, n_edges + 6 <= cap_edges (even), n_triplets <= cap_triplets; o_N has
 * cap_molecules + 1 entries.  Sortedness of batch_seg / id3_reduce_ca is validated on the device by gemnet_csr_from_sorted.
 * ---------------------------------------------------------------------------------------------- */
int gemnet_pad_batch_host(const int64_t* Z, const float* R, const int64_t* batch_seg, const int64_t* N,
                          const int64_t* id_c, const int64_t* id_a, const int64_t* id_swap, const int64_t* id_undir,
                          const int64_t* id3_expand_ba, const int64_t* id3_reduce_ca, int64_t n_atoms, int64_t n_edges,
                          int64_t n_triplets, int64_t n_molecules, int64_t cap_atoms, int64_t cap_edges,
                          int64_t cap_triplets, int64_t cap_molecules, int32_t* o_Z, float* o_R, int32_t* o_batch_seg,
                          int32_t* o_N, int32_t* o_id_c, int32_t* o_id_a, int32_t* o_id_swap, int32_t* o_id_undir,
                          int32_t* o_id3_expand_ba, int32_t* o_id3_reduce_ca);

#ifdef __cplusplus
}
#endif
#endif /* GEMNET_B200_H */
