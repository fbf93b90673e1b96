// kernels.cuh -- argument blocks and launchers shared by kernels.cu (device code) and api.cu (host).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "lattice.cuh"

namespace luma {

struct StepArgs
{
	const double *fin;        // lattice read by this step   [Q][stride]
	double *fout;             // lattice written by this step
	const uint32_t *cw;       // cell words [cells]
	double *rho;              // [cells]
	double *u;                // SoA [D][stride]
	long long stride;         // elements between populations (>= cells, multiple of 16)
	long long off_pull[27];   // BYTE offset 8*(v*stride - (cx*M*K + cy*K + cz)): where population v is pulled from (no wrap, no bounce-back)
	int P, M, K;              // local planes (incl. ghost planes when !wrap_x), rows, columns
	unsigned MK;              // M*K
	int wrap_x;               // 1: periodic wrap inside the array (single rank); 0: ghost planes 0 and P-1
	int p0, pstep;            // plane handled by blockIdx.y: p0 + blockIdx.y * pstep
	int write_macro;          // store rho,u of fluid sites (last step of a call)
	int use_tma;              // measured variants of k_step: 1 = k_step_tma (loads staged through shared memory by bulk copies),
	                          // 2 = k_step_v2 (two sites per thread, 128-bit accesses); 0 = k_step
	int rest_only;            // k_step: only the sites k_step_v2 left alone (its second launch)
	int fill_holes;           // k_step: never-updated sites that share a 32-byte sector with an updated site join its stores (see step_site)
	double omega;
	double tau;               // 1.0 / omega
	double smag_coef;         // 2.0*L_SQRT2*SQ(L_CSMAG)*L_RHOIN*SQ(cs)*SQ(cs)          optimised.cpp:752
	double lam[4];            // (1 - 0.5*omega) * (w/(cs*cs)) per weight class            optimised.cpp:962
	double kbc_beta_m1;       // 2.0 / omega                                               optimised.cpp:1279
	double kbc_inv_beta;      // 1.0 / kbc_beta_m1                                         optimised.cpp:1293
	double Fg;                // force_xyz along L_GRAVITY_DIRECTION (uniform: rho_init * gravity; the other components are 0) init_grids.cpp:296
	double hFg;               // 0.5 * Fg
	LbmConst C;
	// boundary sites
	const long long *bc_list; // local site ids of the sites k_bc handles (classes 2, 3, 4), ascending
	const int *bc_extra;      // per list entry: extra advances of the time averages (see tavg_update), or nullptr
	int n_bc;
	const double *uin;        // [3][M] ux_in, uy_in, uz_in
	double ramp;              // getVelocityRampCoefficient((t+1)*dt)   regularised BC, optimised.cpp:331
	double ramp_t;            // getVelocityRampCoefficient(t*dt)       forced-equilibrium BC, optimised.cpp:258
	double rho_out;
	// per-link handling of class-4 sites (and of regularised sites when `general` is set)
	const uint8_t *types;     // eType per cell
	const uint32_t *bcdesc;   // wall descriptors per cell (read by lattices whose cell word has no room for them)
	int general;              // the grid has eSlip / eExtrapolateRight / forced-equilibrium sources
	int regularised;          // L_REGULARISED_BOUNDARIES
	int velramp_on;           // L_VELOCITY_RAMP defined: forced-equilibrium sources take u = u_in[j]*ramp_t
	// fused halo exchange (k_step_faces, k_bc): the ring neighbours' copies of `fout`, peer-mapped over NVLink;
	// [0] = left (rank-1), [1] = right (rank+1); nullptr = off (k_halo_push copies the planes afterwards instead)
	double *peer_f[2];
	long long peer_stride[2];
	int peer_P[2];
	// time-averaged statistics (L_COMPUTE_TIME_AVERAGED_QUANTITIES, optimised.cpp:895-917)
	double *tav;              // SoA [1 + D + 3D-3][stride]: rho, u_p, u_p*u_q (p <= q) or nullptr
	double t_now, t_next;     // (double)t and (double)(t + 1) of this step
};

// sites whose stored velocity the forced-equilibrium BC overwrites (optimised.cpp:259-263)
struct VelSrcArgs
{
	const long long *list;    // local ids of the non-regularised eVelocity sites on owned planes
	int n;
	const uint8_t *types;
	const uint32_t *bcdesc;   // wall descriptors (slip destinations)
	double *u;
	long long stride;
	const double *uin;
	double ramp_t;
	int P, M, K, N;
	int wrap_x;
	int x_first;              // global x of local plane 0
};

// ---- device-initiated halo exchange over NVLink peer memory (kernels.cu k_halo_push / k_halo_wait) ----
struct HaloPushArgs
{
	const double *src[18];    // local: one population plane of the lattice just written (M*K doubles each)
	double *dst[18];          // the same population's ghost plane in the neighbour's lattice (peer-mapped)
	int nmsg;
	long long count;          // M*K
	unsigned long long *peer_flag[2];   // the neighbours' arrival flags for data coming from this rank
	unsigned long long *seq;  // this rank's exchange number (device memory; advanced by the publishing kernel)
	unsigned int *done;       // local block counter (zero between launches)
};

struct GeomArgs
{
	const uint8_t *types;     // eType per cell [cells]
	const uint32_t *bcdesc;   // cw_pack_bc() bits per cell or nullptr
	uint32_t *cw;
	int P, M, K;
	int wrap_x;
	int p_begin, p_end;       // planes that get a cell word (owned planes)
	int regularised;
};

struct SynthArgs
{
	uint8_t *types;
	uint32_t *bcdesc;
	double *f0, *f1;
	double *rho, *u;
	long long stride;
	int P, M, K;
	int N;                    // global x size
	int x_first;              // global x index of local plane 0 (may be -1 / wrap)
	int wall_type[6];
	int wall_cells[6];
	double uin_uniform[3];
	const double *uin;        // [3][M] profiles (filled on host)
	double ramp0;
	double rhoin;
	int no_flow;
	int has_box;
	int box[6];
	LbmConst C;
};

// coll: 0 BGK, 1 BGK + Smagorinsky, 2 KBC (D2Q9 and D3Q27 only; D3Q27 always); force: 0 none, 1 + L_GRAVITY_DIRECTION
template <class L> void launch_step(const StepArgs &a, int coll, int force, int nplanes, cudaStream_t s, int64_t *launches);
template <class L> void launch_bc(const StepArgs &a, int coll, int force, cudaStream_t s, int64_t *launches);
// k_step on the slab's face planes that also stores the outgoing populations into the neighbours' ghost planes (a.peer_f)
template <class L> void launch_step_faces(const StepArgs &a, int coll, int force, int nplanes, cudaStream_t s, int64_t *launches);
template <class L> void launch_velsrc(const VelSrcArgs &a, cudaStream_t s, int64_t *launches);
void launch_halo_push(const HaloPushArgs &a, cudaStream_t s);
void launch_halo_wait(const unsigned long long *flags, const unsigned long long *seq, int *timed_out, cudaStream_t s);
void launch_halo_publish(unsigned long long *left_flag, unsigned long long *right_flag, unsigned long long *seq, cudaStream_t s);
void launch_force_general(uint32_t *cw, const long long *ids, int n, int class_shift, cudaStream_t s);
void launch_scatter_u32(uint32_t *out, const long long *ids, const uint32_t *vals, int n, cudaStream_t s);
// *flag |= 1 if any cell word in [first, first + n) has one of the bits in `mask`
void launch_any_bits(const uint32_t *cw, long long first, long long n, uint32_t mask, int *flag, cudaStream_t s);
template <class L> void launch_cell_words(const GeomArgs &g, cudaStream_t s);
template <class L> void launch_synthetic(const SynthArgs &a, cudaStream_t s);
// f = feq(rho, u) on sites [first, first + n): LBM_initGrid's population initialisation (init_grids.cpp:310-333)
template <class L> void launch_feq_init(const double *rho, const double *u, double *f, long long stride, long long first, long long n,
	const LbmConst &C, cudaStream_t s);
template <class L> void launch_aos_to_soa(const double *aos, double *soa, long long stride, long long first_cell, long long ncells, cudaStream_t s);
template <class L> void launch_soa_to_aos(const double *soa, double *aos, long long stride, long long first_cell, long long ncells, cudaStream_t s);
// vectors of ncomp = 2, 3 or 6 components per site (u, ui_timeav, uiuj_timeav)
void launch_u_aos_to_soa(const double *aos, double *soa, long long stride, int ncomp, long long first_cell, long long ncells, cudaStream_t s);
void launch_u_soa_to_aos(const double *soa, double *aos, long long stride, int ncomp, long long first_cell, long long ncells, cudaStream_t s);
void launch_types_from_i32(const int32_t *in, uint8_t *out, long long n, cudaStream_t s);
void launch_types_to_i32(const uint8_t *in, int32_t *out, long long n, cudaStream_t s);
void launch_selftest_div(const LbmConst &C, unsigned long long seed, long long n, unsigned long long *mismatches, cudaStream_t s);
// momentum exchange: per-block partial sums [nblocks][3]; returns nblocks
template <class L> int launch_momex(const double *f_prev, const uint8_t *types, long long stride, int P, int M, int K,
	int p_begin, int p_end, int x_first, int N, double *partials, int max_blocks, cudaStream_t s);

}  // namespace luma
