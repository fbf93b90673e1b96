// api.cu -- the C ABI of include/luma_b200.h: handle, device state, upload/download, the step loop
// (GridObj::LBM_multi_opt, src/GridObj_ops_lbm_optimised.cpp:36-193) and the slab halo exchange
// that replaces MpiManager::mpi_communicate (src/MpiManager.cpp:631-815).
//
// There is no CPU fallback anywhere in this file: every entry point that touches state needs a
// CUDA device and fails with LUMA_B200_ECUDA otherwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <dlfcn.h>
#include <cuda_runtime.h>
#include <nccl.h>

#include "../../include/luma_b200.h"
#include "kernels.cuh"

using namespace luma;

// ---- NCCL is bound at run time (dlopen), so single-GPU users need no NCCL at all and a host
//      process that already carries an NCCL (e.g. torch's bundled one) shares it ----
struct NcclApi
{
	void *lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool load(std::string &err)
	{
		if (lib) return true;
		const char *names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
		if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define LUMA_SYM(field, name) *(void **)(&field) = dlsym(lib, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
		LUMA_SYM(GetUniqueId, "ncclGetUniqueId")
		LUMA_SYM(CommInitRank, "ncclCommInitRank")
		LUMA_SYM(CommDestroy, "ncclCommDestroy")
		LUMA_SYM(Send, "ncclSend")
		LUMA_SYM(Recv, "ncclRecv")
		LUMA_SYM(GroupStart, "ncclGroupStart")
		LUMA_SYM(GroupEnd, "ncclGroupEnd")
		LUMA_SYM(GetErrorString, "ncclGetErrorString")
#undef LUMA_SYM
		return true;
	}
};
static NcclApi g_nccl;

// what a rank publishes so that its ring neighbours can store into its ghost planes (luma_b200_p2p_export)
struct P2PBlob
{
	uint32_t magic;
	int32_t rank, device, P;
	long long stride, MK;
	cudaIpcMemHandle_t f[2], flags;
};
static_assert(sizeof(P2PBlob) <= LUMA_B200_P2P_BLOB_BYTES, "P2P blob fits its ABI size");

struct PeerMap
{
	void *base[3] = { nullptr, nullptr, nullptr };   // mapped f[0], f[1], flags of the neighbour (owned by this map)
	double *f[2] = { nullptr, nullptr };
	unsigned long long *flags = nullptr;
	long long stride = 0;
	int P = 0;
};

struct GraphSlot
{
	cudaGraphExec_t exec = nullptr;
	double omega = 0.0;
	int n_bc = -1;
	int epoch = -1;
	int64_t nodes = 0;
};

struct luma_b200
{
	LumaCaseParams p;
	int Q = 0, D = 0;
	int ghost = 0;              // 1 when nranks > 1: local planes 0 and P-1 are ghost planes
	int P = 0;                  // local planes = x_count + 2*ghost
	long long MK = 0, cells = 0, stride = 0;
	double *f[2] = { nullptr, nullptr };
	int cur = 0;
	uint32_t *cw = nullptr;
	uint32_t *bcdesc = nullptr;
	uint8_t *types = nullptr;
	double *rho = nullptr, *u = nullptr, *uin = nullptr;
	long long *bc_list = nullptr;   // sites k_bc handles (classes 2, 3, 4), ascending
	int *bc_extra = nullptr;        // per entry: extra advances of the time averages (reference quirk), or null
	int n_bc = 0;
	long long *vel_list = nullptr;  // forced-equilibrium eVelocity sites whose stored u follows the ramp
	int n_vel = 0;
	bool general = false;           // the grid holds eSlip / eExtrapolateRight / forced-equilibrium sources
	double *tav = nullptr;          // time-averaged statistics, SoA [1 + D + 3D-3][stride]
	void *staging = nullptr;
	size_t staging_bytes = 0;
	uint8_t *types_host = nullptr;  // pinned host mirror of `types` (geometry finalisation)
	size_t types_host_bytes = 0;
	double *momex_dev = nullptr;
	cudaStream_t s_main = nullptr, s_comm = nullptr, s_copy = nullptr;
	cudaEvent_t ev_edge = nullptr, ev_comm = nullptr, ev_int = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_snap = nullptr, ev_copied = nullptr,
		ev_fork = nullptr, ev_join = nullptr;
	double *snap = nullptr;         // snapshot of rho / u (AoS) / f (AoS) feeding an asynchronous download
	size_t snap_bytes = 0;
	bool copy_pending = false;
	ncclComm_t comm = nullptr;
	LbmConst C;
	double omega = 0.0, nu = 0.0;       // GridObj::omega, ::nu after the steps accepted so far (host arithmetic)
	int t = 0;                          // GridObj::t: steps accepted by luma_b200_step (upload's t + ...)
	int t_enq = 0;                      // time level reached by the work ENQUEUED so far; t - t_enq steps are deferred
	double omega_enq = 0.0;             // omega of the last enqueued step (Reynolds ramp)
	bool timing_open = false;           // ev_t0 recorded, ev_t1 not yet: steps have been enqueued since the last read point
	int timed_steps = 0;                // steps enqueued since ev_t0
	bool stream_joined = true;          // s_main has waited for everything issued on s_comm
	bool stats_dirty = false;           // a closed timing window (ev_t0 .. ev_t1) has not been read yet
	int window_steps = 0;               // steps inside that window
	bool have_state = false;
	bool stepped = false;
	LumaStats st;
	std::string err;
	// device-initiated halo exchange (NVLink peer stores); falls back to NCCL send/recv when not attached
	unsigned long long *flags = nullptr;   // [0] exchange number that arrived from the left neighbour, [1] from the right; [4]: block counter of
	                                       // k_halo_push; [5]: this rank's own exchange number (advanced on the device: the launches are replayable)
	int *halo_timeout = nullptr;           // set by k_halo_wait when a neighbour never showed up (device view of a mapped host word)
	volatile int *halo_timeout_host = nullptr;   // the same word as the host reads it: polled without any stream traffic
	PeerMap peer[2];                       // 0 = left (rank-1), 1 = right (rank+1); peer[1] aliases peer[0] when nranks == 2
	bool p2p = false;
	bool fused = false;                    // the face kernels store into the neighbours' ghost planes themselves (default with peer stores; LUMA_B200_FUSED_HALO=0: copy kernel)
	GraphSlot graphs[2];            // captured batches of graph_steps steps, one per lattice parity
	int graph_steps = 0;            // 0 = never use graphs
	bool graph_slabs = true;        // batches on slabs too (device-initiated exchange only); LUMA_B200_GRAPH_SLABS=0 turns them off
	int geometry_epoch = 0;         // bumped by upload / init_synthetic
	bool fill_holes = true;         // solid sites at wall-bounded row ends join the stores of their sector, link sites take the select
	                                // sequence: decided per geometry (walls or bodies bounded in z); LUMA_B200_FILL=0 / 1 forces it off / on
	int fill_env = -1;              // -1 automatic, 0 off, 1 on
	int use_tma = 0;                // LUMA_B200_TMA=1 / LUMA_B200_V2=1 at create: measured variants of k_step (profiles/r02_variants.txt)
	bool profiling = false;
	std::vector<cudaEvent_t> prof_ev;   // pairs (start, stop) recorded during the current step call
	size_t prof_used = 0;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return LUMA_B200_ECUDA; } } while (0)
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
	h->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return LUMA_B200_ENCCL; } } while (0)
#define FAIL(code, msg) do { h->err = (msg); return (code); } while (0)

static const double LUMA_PI = 3.14159265358979323846;     // L_PI, inc/stdafx.h:114
static const double LUMA_SQRT2 = 1.4142135623730950488016887242097;   // L_SQRT2, inc/stdafx.h:113

// GridUtils::getVelocityRampCoefficient, src/GridUtils.cpp:1808-1816
static double velocity_ramp_coef(const LumaCaseParams &p, double t)
{
	if (p.velocity_ramp_on && t <= p.velocity_ramp) return (1.0 - cos(LUMA_PI * t / p.velocity_ramp)) / 2.0;
	return 1.0;
}
// GridUtils::getReynoldsRampCoefficient, src/GridUtils.cpp:1825-1833
static double reynolds_ramp_coef(const LumaCaseParams &p, double t)
{
	if (p.reynolds_ramp_on && t <= p.reynolds_ramp) return 1.0 - cos(LUMA_PI * t / p.reynolds_ramp);
	return 1.0;
}

static void make_constants(LbmConst &C, int Q)
{
	const volatile double three = 3.0, one = 1.0;     // volatile: evaluate at run time like the reference (src/stdafx.cpp:153)
	const double cs = one / sqrt(three);
	C.cs2 = cs * cs;
	C.inv_cs2 = 1.0 / C.cs2;
	C.den = (2.0 * C.cs2) * C.cs2;
	C.inv_den = 1.0 / C.den;
	C.k1 = 1.0 - C.cs2;
	C.k0 = 0.0 - C.cs2;
	C.w[3] = 0.0;
	if (Q == 27) { C.w[0] = 2.0 / 27.0; C.w[1] = 1.0 / 54.0; C.w[2] = 1.0 / 216.0; C.w[3] = 8.0 / 27.0; }   // src/stdafx.cpp:130-136
	else if (Q == 19) { C.w[0] = 1.0 / 18.0; C.w[1] = 1.0 / 36.0; C.w[2] = 1.0 / 3.0; }  // :140-143
	else { C.w[0] = 1.0 / 9.0; C.w[1] = 1.0 / 36.0; C.w[2] = 4.0 / 9.0; }                // :147-148
	for (int k = 0; k < 4; ++k) C.wden[k] = C.w[k] / C.den;
}

// next (start or stop) event of the per-kernel profile, or nullptr when the runtime cannot create one: the caller then
// skips the pair and the launch is simply not part of the profile
static cudaEvent_t prof_event(luma_b200_t *h)
{
	if (h->prof_used == h->prof_ev.size())
	{
		cudaEvent_t e = nullptr;
		if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		h->prof_ev.push_back(e);
	}
	return h->prof_ev[h->prof_used++];
}

// lattice dispatch: runs CALL with L bound to the lattice of a handle (L_NUM_VELS 9, 19 or 27)
#define LAT(Q_, CALL) do { if ((Q_) == 19) { using L = D3Q19; CALL; } else if ((Q_) == 27) { using L = D3Q27; CALL; } \
	else { using L = D2Q9; CALL; } } while (0)

template <class L>
static void main_kernel(luma_b200_t *h, const StepArgs &a, int coll, int force, int nplanes)
{
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (h->profiling && nplanes > 0)
	{
		const size_t used = h->prof_used;
		e0 = prof_event(h);
		e1 = e0 ? prof_event(h) : nullptr;
		if (!e1) { h->prof_used = used; e0 = nullptr; }      // events come in (start, stop) pairs or not at all
	}
	if (e0) cudaEventRecord(e0, h->s_main);
	launch_step<L>(a, coll, force, nplanes, h->s_main, &h->st.kernel_launches);
	if (e1)
	{
		cudaEventRecord(e1, h->s_main);
		h->st.step_kernel_launches++;
		h->st.step_kernel_cells += (long long)nplanes * h->MK;
	}
}

static int halo_health(luma_b200_t *h)
{
	if (h->halo_timeout_host && *h->halo_timeout_host)
		FAIL(LUMA_B200_ENCCL, "halo exchange: a ring neighbour did not deliver its populations within 20 s");
	return LUMA_B200_OK;
}

extern "C" {

int luma_b200_abi_version(void) { return LUMA_B200_ABI_VERSION; }

const char *luma_b200_strerror(int code)
{
	switch (code)
	{
	case LUMA_B200_OK: return "ok";
	case LUMA_B200_EINVAL: return "invalid argument or inconsistent case description";
	case LUMA_B200_ECUDA: return "CUDA runtime error";
	case LUMA_B200_ENCCL: return "NCCL error";
	case LUMA_B200_ENOMEM: return "out of memory";
	case LUMA_B200_EUNSUPPORTED: return "feature outside the level-0 BGK/Smagorinsky/KBC path";
	case LUMA_B200_ESTATE: return "call out of order";
	case LUMA_B200_EBC_NOT_WALL: return "Trying to apply a regularised BC on a site not within a wall.";
	case LUMA_B200_EBC_PRESSURE_EDGE: return "Pressure BC cannot be applied to a corner or an edge.";
	case LUMA_B200_EBC_OFFGRID: return "Extrapolation site off grid";
	default: return "unknown luma_b200 status";
	}
}

const char *luma_b200_last_error(luma_b200_t *h) { return h ? h->err.c_str() : "null handle"; }

void luma_b200_default_params(LumaCaseParams *p)
{
	if (!p) return;
	memset(p, 0, sizeof(*p));
	p->struct_size = (uint32_t)sizeof(LumaCaseParams);
	p->dims = 3; p->num_vels = 19;
	p->K = 1;
	p->nranks = 1;
	p->regularised = 1;
	p->csmag = 0.3;
	p->rhoin = 1.0; p->rho_out = 1.0;
	p->omega = 1.0;
	p->re = 1.0;
}

int luma_b200_slab(int32_t N, int32_t nranks, int32_t rank, int32_t *x_offset, int32_t *x_count)
{
	if (N < 1 || nranks < 1 || rank < 0 || rank >= nranks) return LUMA_B200_EINVAL;
	int per = (int)std::ceil((double)N / (double)nranks);
	int last = per - (per * nranks - N);
	if (last <= 0)
	{
		per = (int)std::floor((double)N / (double)nranks);
		last = per - (per * nranks - N);
		if (last <= 0) return LUMA_B200_EINVAL;
	}
	if (per < 1) return LUMA_B200_EINVAL;   /* a rank without planes: the reference would build an empty grid */
	if (x_offset) *x_offset = per * rank;
	if (x_count) *x_count = (rank == nranks - 1) ? last : per;
	return LUMA_B200_OK;
}

static void free_all(luma_b200_t *h)
{
	if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
	cudaFree(h->f[0]); cudaFree(h->f[1]); cudaFree(h->cw); cudaFree(h->bcdesc); cudaFree(h->types);
	cudaFree(h->rho); cudaFree(h->u); cudaFree(h->uin); cudaFree(h->bc_list); cudaFree(h->staging); cudaFree(h->momex_dev);
	cudaFree(h->bc_extra); cudaFree(h->vel_list); cudaFree(h->tav);
	if (h->types_host) cudaFreeHost(h->types_host);
	if (h->ev_edge) cudaEventDestroy(h->ev_edge);
	if (h->ev_comm) cudaEventDestroy(h->ev_comm);
	if (h->ev_t0) cudaEventDestroy(h->ev_t0);
	if (h->ev_t1) cudaEventDestroy(h->ev_t1);
	if (h->ev_snap) cudaEventDestroy(h->ev_snap);
	if (h->ev_fork) cudaEventDestroy(h->ev_fork);
	if (h->ev_join) cudaEventDestroy(h->ev_join);
	if (h->ev_int) cudaEventDestroy(h->ev_int);
	if (h->ev_copied) cudaEventDestroy(h->ev_copied);
	if (h->s_copy) cudaStreamDestroy(h->s_copy);
	cudaFree(h->snap);
	for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
	for (GraphSlot &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
	for (int side = 0; side < 2; ++side)
		for (void *&b : h->peer[side].base) { if (b) cudaIpcCloseMemHandle(b); b = nullptr; }
	cudaFree(h->flags);
	if (h->halo_timeout_host) cudaFreeHost((void *)h->halo_timeout_host);
	if (h->s_main) cudaStreamDestroy(h->s_main);
	if (h->s_comm) cudaStreamDestroy(h->s_comm);
}

int luma_b200_create(luma_b200_t **out, const LumaCaseParams *p)
{
	if (!out || !p) return LUMA_B200_EINVAL;
	*out = nullptr;
	if (p->struct_size != sizeof(LumaCaseParams)) return LUMA_B200_EINVAL;
	// L_NUM_VELS follows from L_DIMS and L_USE_KBC_COLLISION (inc/definitions.h:299-310)
	if (!((p->dims == 3 && p->num_vels == (p->kbc ? 27 : 19)) || (p->dims == 2 && p->num_vels == 9))) return LUMA_B200_EINVAL;
	// "Cannot use regularised boundaries with D3Q27 because of the corner treatment" (src/GridObj_init_grids.cpp:266-270)
	if (p->num_vels == 27 && p->regularised) return LUMA_B200_EINVAL;
	if (p->N < 1 || p->M < 2 || p->K < 1 || (p->dims == 2 && p->K != 1)) return LUMA_B200_EINVAL;
	if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) return LUMA_B200_EINVAL;
	if (p->x_count < 1 || p->x_offset < 0 || p->x_offset + p->x_count > p->N) return LUMA_B200_EINVAL;
	if (p->nranks == 1 && (p->x_offset != 0 || p->x_count != p->N)) return LUMA_B200_EINVAL;
	if (p->gravity_on && (p->gravity_dir < 0 || p->gravity_dir >= p->dims)) return LUMA_B200_EINVAL;
	if (!(p->omega > 0.0)) return LUMA_B200_EINVAL;
	if (!p->bgksmag && !p->reynolds_ramp_on && p->omega >= 2.0) return LUMA_B200_EINVAL;   // init_grids.cpp:353-356

	luma_b200_t *h = new (std::nothrow) luma_b200();
	if (!h) return LUMA_B200_ENOMEM;
	*out = h;       // returned even on failure so that luma_b200_last_error() can be read; destroy it
	h->p = *p;
	h->Q = p->num_vels; h->D = p->dims;
	h->ghost = (p->nranks > 1) ? 1 : 0;
	h->P = p->x_count + 2 * h->ghost;
	h->MK = (long long)p->M * p->K;
	h->cells = (long long)h->P * h->MK;
	h->stride = (h->cells + 15) / 16 * 16;
	{
		// tuning knob (profiles/r02_variants.txt): extra elements between the population arrays, a multiple of 16, so that
		// the Q read streams and Q write streams of the step do not sit at power-of-two distances from each other
		const char *pv = getenv("LUMA_B200_STRIDE_PAD");
		if (pv && *pv && atoll(pv) > 0) h->stride += (atoll(pv) + 15) / 16 * 16;
	}
	h->omega = p->omega;
	h->t = p->t;
	memset(&h->st, 0, sizeof(h->st));
	h->st.cells = (long long)p->x_count * h->MK;
	make_constants(h->C, h->Q);
	{
		// CUDA-graph batches: by default for grids small enough to be launch-bound; LUMA_B200_GRAPH_STEPS=0 turns
		// them off, LUMA_B200_GRAPH_CELLS moves the size limit
		const char *gsv = getenv("LUMA_B200_GRAPH_STEPS"), *gcv = getenv("LUMA_B200_GRAPH_CELLS");
		const long long limit = (gcv && *gcv) ? atoll(gcv) : (4LL << 20);
		int gsteps = (gsv && *gsv) ? atoi(gsv) : 16;
		if (gsteps < 2 || h->cells > limit) gsteps = 0;
		h->graph_steps = gsteps & ~1;
		const char *sv = getenv("LUMA_B200_GRAPH_SLABS");
		h->graph_slabs = !(sv && *sv && atoi(sv) == 0);
	}
	h->nu = (1.0 / h->omega - 0.5) * h->C.cs2;
	{
		const char *tv = getenv("LUMA_B200_TMA"), *v2 = getenv("LUMA_B200_V2");
		h->use_tma = (tv && *tv && atoi(tv) != 0) ? 1 : ((v2 && *v2 && atoi(v2) != 0) ? 2 : 0);
		const char *fv = getenv("LUMA_B200_FILL");
		h->fill_env = (fv && *fv) ? (atoi(fv) != 0 ? 1 : 0) : -1;
		h->fill_holes = h->fill_env != 0;
	}

	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	if (ndev < 1 || p->device < 0 || p->device >= ndev) FAIL(LUMA_B200_ECUDA, "no such CUDA device");
	CK(cudaSetDevice(p->device));
	CK(cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
	{
		// the exchange must not queue behind the interior kernel's 65k blocks: highest priority for its stream
		int lo = 0, hi = 0;
		CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		CK(cudaStreamCreateWithPriority(&h->s_comm, cudaStreamNonBlocking, hi));
	}
	CK(cudaEventCreateWithFlags(&h->ev_edge, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
	CK(cudaEventCreate(&h->ev_t0));
	CK(cudaEventCreate(&h->ev_t1));
	CK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_int, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
	const size_t fbytes = (size_t)h->stride * h->Q * sizeof(double);
	cudaError_t e = cudaMalloc(&h->f[0], fbytes);
	if (e == cudaSuccess) e = cudaMalloc(&h->f[1], fbytes);
	if (e == cudaSuccess) e = cudaMalloc(&h->cw, (size_t)h->cells * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMalloc(&h->bcdesc, (size_t)h->cells * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaMalloc(&h->types, (size_t)h->cells);
	if (e == cudaSuccess) e = cudaMalloc(&h->rho, (size_t)h->stride * sizeof(double));
	if (e == cudaSuccess) e = cudaMalloc(&h->u, (size_t)h->stride * h->D * sizeof(double));
	if (e == cudaSuccess) e = cudaMalloc(&h->uin, (size_t)3 * p->M * sizeof(double));
	if (e == cudaSuccess) e = cudaMalloc(&h->momex_dev, (size_t)3 * 4096 * sizeof(double));
	if (e == cudaSuccess && h->ghost) e = cudaMalloc(&h->flags, 64);
	if (e == cudaSuccess && h->ghost)
	{
		// the time-out word lives in mapped host memory: k_halo_wait stores to it with system scope, the host polls it
		// without touching a stream (luma_b200_step never synchronises)
		void *hp = nullptr, *dp = nullptr;
		e = cudaHostAlloc(&hp, sizeof(int), cudaHostAllocMapped);
		if (e == cudaSuccess) { *(int *)hp = 0; h->halo_timeout_host = (volatile int *)hp; e = cudaHostGetDevicePointer(&dp, hp, 0); }
		if (e == cudaSuccess) h->halo_timeout = (int *)dp;
	}
	const size_t tav_bytes = (size_t)h->stride * (1 + h->D + 3 * h->D - 3) * sizeof(double);
	if (e == cudaSuccess && p->time_averaged) e = cudaMalloc(&h->tav, tav_bytes);
	if (e != cudaSuccess) { h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return LUMA_B200_ENOMEM; }
	CK(cudaMemsetAsync(h->cw, 0, (size_t)h->cells * sizeof(uint32_t), h->s_main));
	CK(cudaMemsetAsync(h->bcdesc, 0, (size_t)h->cells * sizeof(uint32_t), h->s_main));
	CK(cudaMemsetAsync(h->uin, 0, (size_t)3 * p->M * sizeof(double), h->s_main));
	if (h->tav) CK(cudaMemsetAsync(h->tav, 0, tav_bytes, h->s_main));      // init_grids.cpp:304-306
	if (h->flags) CK(cudaMemsetAsync(h->flags, 0, 64, h->s_main));
	// both lattices start as zeros: ghost planes and never-updated sites hold the same, defined bits whichever way the
	// state arrives (upload or init_synthetic) -- nothing reads them, but nothing should depend on cudaMalloc's leftovers
	CK(cudaMemsetAsync(h->f[0], 0, fbytes, h->s_main));
	CK(cudaMemsetAsync(h->f[1], 0, fbytes, h->s_main));
	CK(cudaMemsetAsync(h->rho, 0, (size_t)h->stride * sizeof(double), h->s_main));
	CK(cudaMemsetAsync(h->u, 0, (size_t)h->stride * h->D * sizeof(double), h->s_main));
	CK(cudaMemsetAsync(h->types, 0, (size_t)h->cells, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	h->t_enq = h->t; h->omega_enq = h->omega;
	return LUMA_B200_OK;
}

void luma_b200_destroy(luma_b200_t *h)
{
	if (!h) return;
	cudaSetDevice(h->p.device);
	cudaDeviceSynchronize();
	free_all(h);
	delete h;
}

int luma_b200_comm_unique_id(void *unique_id_128)
{
	std::string err;
	if (!unique_id_128) return LUMA_B200_EINVAL;
	if (!g_nccl.load(err)) return LUMA_B200_ENCCL;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) return LUMA_B200_ENCCL;
	memcpy(unique_id_128, &id, 128);
	return LUMA_B200_OK;
}

int luma_b200_comm_init(luma_b200_t *h, const void *unique_id_128)
{
	if (!h || !unique_id_128) return LUMA_B200_EINVAL;
	if (h->p.nranks < 2) FAIL(LUMA_B200_ESTATE, "comm_init on a single-rank handle");
	if (!g_nccl.load(h->err)) return LUMA_B200_ENCCL;
	CK(cudaSetDevice(h->p.device));
	ncclUniqueId id;
	memcpy(&id, unique_id_128, 128);
	NK(g_nccl.CommInitRank(&h->comm, h->p.nranks, id, h->p.rank));
	return LUMA_B200_OK;
}

int luma_b200_p2p_export(luma_b200_t *h, void *blob)
{
	if (!h || !blob) return LUMA_B200_EINVAL;
	if (!h->ghost) FAIL(LUMA_B200_ESTATE, "p2p_export on a single-rank handle");
	CK(cudaSetDevice(h->p.device));
	P2PBlob b;
	memset(&b, 0, sizeof(b));
	b.magic = 0x4C423230u; b.rank = h->p.rank; b.device = h->p.device; b.P = h->P;
	b.stride = h->stride; b.MK = h->MK;
	CK(cudaIpcGetMemHandle(&b.f[0], h->f[0]));
	CK(cudaIpcGetMemHandle(&b.f[1], h->f[1]));
	CK(cudaIpcGetMemHandle(&b.flags, h->flags));
	memset(blob, 0, LUMA_B200_P2P_BLOB_BYTES);
	memcpy(blob, &b, sizeof(b));
	return LUMA_B200_OK;
}

int luma_b200_p2p_attach(luma_b200_t *h, const void *left_blob, const void *right_blob)
{
	if (!h || !left_blob || !right_blob) return LUMA_B200_EINVAL;
	if (!h->ghost) FAIL(LUMA_B200_ESTATE, "p2p_attach on a single-rank handle");
	if (h->p2p) FAIL(LUMA_B200_ESTATE, "p2p_attach called twice");
	CK(cudaSetDevice(h->p.device));
	const int n = h->p.nranks, want[2] = { (h->p.rank - 1 + n) % n, (h->p.rank + 1) % n };
	const void *blobs[2] = { left_blob, right_blob };
	for (int side = 0; side < 2; ++side)
	{
		P2PBlob b;
		memcpy(&b, blobs[side], sizeof(b));
		if (b.magic != 0x4C423230u || b.rank != want[side] || b.MK != h->MK)
			FAIL(LUMA_B200_EINVAL, "p2p_attach: blob does not come from the ring neighbour");
		PeerMap &pm = h->peer[side];
		pm.stride = b.stride; pm.P = b.P;
		if (side == 1 && n == 2)
		{
			// both neighbours are the same GPU: one mapping (an allocation can be opened once per process)
			pm.f[0] = h->peer[0].f[0]; pm.f[1] = h->peer[0].f[1]; pm.flags = h->peer[0].flags;
			continue;
		}
		const cudaIpcMemHandle_t *hd[3] = { &b.f[0], &b.f[1], &b.flags };
		for (int k = 0; k < 3; ++k)
		{
			const cudaError_t e = cudaIpcOpenMemHandle(&pm.base[k], *hd[k], cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess)
			{
				cudaGetLastError();
				h->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) + " (no peer access between the two GPUs? use the NCCL exchange)";
				return LUMA_B200_ECUDA;
			}
		}
		pm.f[0] = (double *)pm.base[0]; pm.f[1] = (double *)pm.base[1]; pm.flags = (unsigned long long *)pm.base[2];
	}
	h->p2p = true;
	for (GraphSlot &g : h->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
	{
		// the peer stores are part of the face kernels' epilogue (default: measured best at every slab size,
		// profiles/r02_halo_transports_n2.txt); LUMA_B200_FUSED_HALO=0 keeps the separate copy kernel k_halo_push
		const char *fv = getenv("LUMA_B200_FUSED_HALO");
		h->fused = !(fv && *fv && atoi(fv) == 0);
	}
	return LUMA_B200_OK;
}

static int ensure_staging(luma_b200_t *h, size_t bytes)
{
	if (h->staging_bytes >= bytes) return LUMA_B200_OK;
	if (h->staging) { cudaFree(h->staging); h->staging = nullptr; h->staging_bytes = 0; }
	if (cudaMalloc(&h->staging, bytes) != cudaSuccess) { cudaGetLastError(); FAIL(LUMA_B200_ENOMEM, "staging buffer"); }
	h->staging_bytes = bytes;
	return LUMA_B200_OK;
}

// populations with c_x = +1 travel to the +x neighbour, c_x = -1 to the -x neighbour
static int edge_pops(int Q, int sign, int out[9])
{
	int n = 0;
	for (int v = 0; v < Q; ++v)
	{
		int cx = 0;
		LAT(Q, cx = L::c(v, 0));
		if (cx == sign) out[n++] = v;
	}
	return n;
}

// the exchange step: only the populations that cross the slab face, straight from / into the SoA
// lattice (each is one contiguous M*K run), ring topology (MPI_Cart_create periodic, MpiManager.cpp:112-136)
static int build_halo_plan(const LumaCaseParams &p, std::vector<LumaHaloMsg> &plan)
{
	plan.clear();
	if (p.nranks < 2) return 0;
	const int n = p.nranks, right = (p.rank + 1) % n, left = (p.rank - 1 + n) % n;
	int plus[9], minus[9];
	const int np = edge_pops(p.num_vels, +1, plus), nm = edge_pops(p.num_vels, -1, minus);
	const int P = p.x_count + 2;
	for (int a = 0; a < np; ++a) plan.push_back({ 1, right, plus[a], P - 2 });
	for (int a = 0; a < np; ++a) plan.push_back({ 0, left, plus[a], 0 });
	for (int a = 0; a < nm; ++a) plan.push_back({ 1, left, minus[a], 1 });
	for (int a = 0; a < nm; ++a) plan.push_back({ 0, right, minus[a], P - 1 });
	return (int)plan.size();
}

// the same plan executed by this GPU alone: every send becomes stores into the receiver's ghost plane
// (lattice index `li` on both sides: the ranks step in lockstep), receives become a wait on the arrival flags
static int exchange_populations_p2p(luma_b200_t *h, int li, cudaStream_t s, bool already_stored)
{
	if (already_stored)
	{
		// fused exchange: k_step_faces / k_bc of this step wrote the neighbours' ghost planes; only the arrival flags remain
		launch_halo_publish(h->peer[0].flags + 1, h->peer[1].flags + 0, h->flags + 5, s);
		launch_halo_wait(h->flags, h->flags + 5, h->halo_timeout, s);
		h->st.kernel_launches += 2;
		CK(cudaGetLastError());
		return LUMA_B200_OK;
	}
	std::vector<LumaHaloMsg> plan;
	build_halo_plan(h->p, plan);
	HaloPushArgs a;
	memset(&a, 0, sizeof(a));
	for (const LumaHaloMsg &m : plan)
	{
		if (!m.is_send) continue;
		// c_x = +1 populations go to the right neighbour's low ghost plane (0), c_x = -1 populations to the left
		// neighbour's high ghost plane (P_left - 1); with two ranks both neighbours are the same GPU
		int cx = 0;
		LAT(h->Q, cx = L::c(m.pop, 0));
		const bool to_right = cx > 0;
		const PeerMap &pm = h->peer[to_right ? 1 : 0];
		const int dst_plane = to_right ? 0 : pm.P - 1;
		a.src[a.nmsg] = h->f[li] + (long long)m.pop * h->stride + (long long)m.plane * h->MK;
		a.dst[a.nmsg] = pm.f[li] + (long long)m.pop * pm.stride + (long long)dst_plane * h->MK;
		++a.nmsg;
	}
	a.count = h->MK;
	a.peer_flag[0] = h->peer[0].flags + 1;      // we are the left neighbour's RIGHT neighbour
	a.peer_flag[1] = h->peer[1].flags + 0;      // and the right neighbour's LEFT neighbour
	a.seq = h->flags + 5;
	a.done = reinterpret_cast<unsigned int *>(h->flags + 4);
	launch_halo_push(a, s);
	launch_halo_wait(h->flags, h->flags + 5, h->halo_timeout, s);
	h->st.kernel_launches += 2;
	CK(cudaGetLastError());
	return LUMA_B200_OK;
}

static int exchange_populations(luma_b200_t *h, double *lat, cudaStream_t s, bool already_stored = false)
{
	if (h->p2p && (lat == h->f[0] || lat == h->f[1])) return exchange_populations_p2p(h, lat == h->f[0] ? 0 : 1, s, already_stored);
	std::vector<LumaHaloMsg> plan;
	build_halo_plan(h->p, plan);
	const size_t cnt = (size_t)h->MK;
	NK(g_nccl.GroupStart());
	for (const LumaHaloMsg &m : plan)
	{
		double *ptr = lat + (long long)m.pop * h->stride + (long long)m.plane * h->MK;
		if (m.is_send) NK(g_nccl.Send(ptr, cnt, ncclFloat64, m.peer, h->comm, s));
		else NK(g_nccl.Recv(ptr, cnt, ncclFloat64, m.peer, h->comm, s));
	}
	NK(g_nccl.GroupEnd());
	return LUMA_B200_OK;
}

// one-off, when the geometry is finalised: the ghost planes' eType, wall descriptors, rho and u (what a
// site on a slab face needs to know about the sites it pulls from in the neighbouring slab)
static int exchange_ghost_planes(luma_b200_t *h, cudaStream_t s)
{
	const int n = h->p.nranks, right = (h->p.rank + 1) % n, left = (h->p.rank - 1 + n) % n;
	const size_t cnt = (size_t)h->MK;
	const long long lo_own = h->MK, hi_own = (long long)(h->P - 2) * h->MK, lo_gh = 0, hi_gh = (long long)(h->P - 1) * h->MK;
	NK(g_nccl.GroupStart());
	NK(g_nccl.Send(h->types + hi_own, cnt, ncclUint8, right, h->comm, s));
	NK(g_nccl.Recv(h->types + lo_gh, cnt, ncclUint8, left, h->comm, s));
	NK(g_nccl.Send(h->types + lo_own, cnt, ncclUint8, left, h->comm, s));
	NK(g_nccl.Recv(h->types + hi_gh, cnt, ncclUint8, right, h->comm, s));
	NK(g_nccl.Send(h->bcdesc + hi_own, cnt, ncclUint32, right, h->comm, s));
	NK(g_nccl.Recv(h->bcdesc + lo_gh, cnt, ncclUint32, left, h->comm, s));
	NK(g_nccl.Send(h->bcdesc + lo_own, cnt, ncclUint32, left, h->comm, s));
	NK(g_nccl.Recv(h->bcdesc + hi_gh, cnt, ncclUint32, right, h->comm, s));
	for (int d = -1; d < h->D; ++d)
	{
		double *q = (d < 0) ? h->rho : h->u + (long long)d * h->stride;
		NK(g_nccl.Send(q + hi_own, cnt, ncclFloat64, right, h->comm, s));
		NK(g_nccl.Recv(q + lo_gh, cnt, ncclFloat64, left, h->comm, s));
		NK(g_nccl.Send(q + lo_own, cnt, ncclFloat64, left, h->comm, s));
		NK(g_nccl.Recv(q + hi_gh, cnt, ncclFloat64, right, h->comm, s));
	}
	NK(g_nccl.GroupEnd());
	return LUMA_B200_OK;
}

static inline int lat_c(int Q, int v, int d) { int c = 0; LAT(Q, c = L::c(v, d)); return c; }

// after h->types (owned planes) and h->bcdesc exist on the device: ghost planes, validation of the
// boundary sites the way the reference would L_ERROR on them, the list of sites k_bc handles, cell words.
// `desc_of(id, plane, j, k)` gives the packed wall descriptor of an OWNED site (the dense device array
// h->bcdesc holds the same values; the host does not download it).
extern "C++" {
template <class DescFn>
static int finalize_geometry(luma_b200_t *h, DescFn desc_of)
{
	const LumaCaseParams &p = h->p;
	if (h->ghost)
	{
		if (!h->comm) FAIL(LUMA_B200_ESTATE, "nranks > 1 needs luma_b200_comm_init before upload/init");
		int rc = exchange_ghost_planes(h, h->s_main);
		if (rc) return rc;
	}
	// host mirror of the eType array: pinned (the copy runs at PCIe speed), kept for later uploads
	if (h->types_host_bytes < (size_t)h->cells + 8)
	{
		if (h->types_host) cudaFreeHost(h->types_host);
		h->types_host = nullptr; h->types_host_bytes = 0;
		if (cudaHostAlloc((void **)&h->types_host, (size_t)h->cells + 8, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); FAIL(LUMA_B200_ENOMEM, "host mirror of the eType array"); }
		h->types_host_bytes = (size_t)h->cells + 8;
	}
	uint8_t *types = h->types_host;
	memset(types + h->cells, 0, 8);
	CK(cudaMemcpyAsync(types, h->types, (size_t)h->cells, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));

	const int P = h->P, M = p.M, K = p.K, Q = h->Q, D = h->D;
	const int pb = h->ghost, pe = P - h->ghost;
	const bool wrap = h->ghost == 0, reg = p.regularised != 0;
	auto site = [&](int pl, int j, int k) { return ((long long)pl * M + j) * K + k; };
	auto never_streamed = [&](uint8_t t) { return t == T_SOLID || t == T_REFINED || (t == T_VELOCITY && !reg); };   // optimised.cpp:91-95

	// The scan over all sites runs on a few host threads, each over a contiguous range of planes; the per-range results are
	// concatenated in plane order (lists stay ascending) and the first failure in plane order is the one reported.
	struct ScanOut
	{
		std::vector<long long> list;                 // sites k_bc handles
		std::vector<long long> forced;               // eFluid sites that must be handled per link (class 4)
		std::vector<std::pair<long long, int>> extra; // (site, extra advances of its time averages per step)
		std::vector<long long> vel;                  // forced-equilibrium eVelocity sites (their stored u follows the ramp)
		bool general = false;
		int rc = LUMA_B200_OK;
		std::string err;
	};
	auto scan_planes = [&](const int pl_begin, const int pl_end, ScanOut &o) -> int
	{
	std::vector<long long> &list = o.list, &forced = o.forced, &vel = o.vel;
	std::vector<std::pair<long long, int>> &extra = o.extra;
	bool &general = o.general;
#define SCAN_FAIL(code, msg) do { o.err = (msg); o.rc = (code); return (code); } while (0)
	for (int pl = pl_begin; pl < pl_end; ++pl)
	{
		const bool owned = pl >= pb && pl < pe;
		const uint8_t *row = types + (size_t)pl * M * K;
		for (int j = 0; j < M; ++j)
			for (int k = 0; k < K; ++k)
			{
				if (K - k >= 8)
				{
					// eight sites at a time: nothing to do where all of them are eSolid (0) or eFluid (1)
					uint64_t w8;
					memcpy(&w8, row + (size_t)j * K + k, 8);
					if ((w8 & 0xFEFEFEFEFEFEFEFEull) == 0) { k += 7; continue; }
				}
				const uint8_t t = row[(size_t)j * K + k];
				if (t == T_SOLID || t == T_FLUID) continue;
				const long long id = site(pl, j, k);
				if (t != T_VELOCITY && t != T_PRESSURE && t != T_SLIP && t != T_EXTRAPOLATE_RIGHT)
					SCAN_FAIL(LUMA_B200_EUNSUPPORTED, "site type " + std::to_string((int)t) + " (refinement/BFL) is outside the level-0 path");

				// sites that pull from an eExtrapolateRight or forced-equilibrium eVelocity site take the per-link path
				if (t == T_EXTRAPOLATE_RIGHT || (t == T_VELOCITY && !reg))
				{
					general = true;
					for (int v = 0; v < Q; ++v)
					{
						int dp = pl + lat_c(Q, v, 0), dj = j + lat_c(Q, v, 1), dk = k + lat_c(Q, v, 2);
						if (wrap) dp = (dp + P) % P;
						if (dp < pb || dp >= pe) continue;
						dj = (dj + M) % M; dk = (dk + K) % K;
						const uint8_t dt = types[(size_t)site(dp, dj, dk)];
						if (never_streamed(dt)) continue;
						if (t == T_EXTRAPOLATE_RIGHT && pl - 2 < pb)
						{
							// optimised.cpp:249-250 reads two planes to the left of the source
							if (h->ghost) SCAN_FAIL(LUMA_B200_EUNSUPPORTED, "an eExtrapolateRight site needs two planes to its left inside the same slab (slab too thin, or the site is reached through the periodic wrap)");
							SCAN_FAIL(LUMA_B200_EBC_OFFGRID, "eExtrapolateRight site within two planes of the low x end: the reference reads off the array");
						}
						if (dt == T_FLUID) forced.push_back(site(dp, dj, dk));
					}
				}
				if (!owned) continue;

				if (t == T_SLIP)
				{
					general = true;
					if ((desc_of(id, pl, j, k) >> CW_EC_SHIFT) == 0)
						SCAN_FAIL(LUMA_B200_EBC_NOT_WALL, "Slip wall not located inside a domain wall region. Not currently supported.");   // optimised.cpp:577
					list.push_back(id);
					continue;
				}
				if (t == T_EXTRAPOLATE_RIGHT) { list.push_back(id); continue; }
				if (!reg)
				{
					// non-regularised build: ePressure sites stream and collide with their stored rho,u; eVelocity sites are skipped
					if (t == T_PRESSURE) { general = true; list.push_back(id); }
					else vel.push_back(id);
					continue;
				}

				// regularised velocity / pressure site (optimised.cpp:313-510)
				const uint32_t d = desc_of(id, pl, j, k);
				const int ec = (int)(d >> CW_EC_SHIFT);
				if (ec == 0) SCAN_FAIL(LUMA_B200_EBC_NOT_WALL, luma_b200_strerror(LUMA_B200_EBC_NOT_WALL));
				if (ec > 1 && t == T_PRESSURE) SCAN_FAIL(LUMA_B200_EBC_PRESSURE_EDGE, luma_b200_strerror(LUMA_B200_EBC_PRESSURE_EDGE));
				if (ec > 1 || t == T_PRESSURE)
				{
					int n[3];
					for (int a = 0; a < 3; ++a) n[a] = (int)((d >> (CW_N_SHIFT + 2 * a)) & 3u) - 1;
					const int ncalls = (t == T_PRESSURE) ? D - 1 : 1;      // one _LBM_updateAndExtrapolate per extrapolated quantity
					for (int m = 1; m <= 2; ++m)
					{
						const int gi = p.x_offset + (pl - h->ghost) + m * n[0], jj = j + m * n[1], kk = k + m * n[2];
						if (gi < 0 || gi >= p.N || jj < 0 || jj >= M || kk < 0 || kk >= K)
							SCAN_FAIL(LUMA_B200_EBC_OFFGRID, luma_b200_strerror(LUMA_B200_EBC_OFFGRID));
						const int pp = pl + m * n[0];
						if (pp < pb || pp >= pe)
							SCAN_FAIL(LUMA_B200_EUNSUPPORTED, "slab too thin: a boundary site extrapolates from a plane owned by another rank");
						const long long idn = site(pp, jj, kk);
						const uint8_t tn = types[(size_t)idn];
						if (tn != T_SOLID && tn != T_FLUID)
							SCAN_FAIL(LUMA_B200_EUNSUPPORTED, "a boundary site extrapolates from another boundary site (loop-order dependent in the reference)");
						if (idn > id)
						{
							// the reference streams + macros this neighbour early (optimised.cpp:1375-1404)
							if (tn == T_SOLID)
								SCAN_FAIL(LUMA_B200_EUNSUPPORTED, "a boundary site extrapolates from an eSolid site with a larger index (the reference streams into that solid site)");
							if (p.time_averaged) { extra.push_back({ idn, ncalls }); forced.push_back(idn); }
						}
					}
				}
				list.push_back(id);
			}
	}
	return LUMA_B200_OK;
#undef SCAN_FAIL
	};
	int nthreads = 1;
	if (h->cells > (4LL << 20))
	{
		nthreads = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
		nthreads = std::min(nthreads, P);
	}
	std::vector<ScanOut> parts((size_t)nthreads);
	if (nthreads == 1) scan_planes(0, P, parts[0]);
	else
	{
		std::vector<std::thread> pool;
		for (int th = 0; th < nthreads; ++th)
			pool.emplace_back([&, th] { scan_planes((int)((long long)P * th / nthreads), (int)((long long)P * (th + 1) / nthreads), parts[(size_t)th]); });
		for (std::thread &t : pool) t.join();
	}
	std::vector<long long> list, forced, vel;
	std::vector<std::pair<long long, int>> extra;
	bool general = false;
	for (ScanOut &o : parts)
	{
		if (o.rc) FAIL(o.rc, o.err);
		list.insert(list.end(), o.list.begin(), o.list.end());
		forced.insert(forced.end(), o.forced.begin(), o.forced.end());
		extra.insert(extra.end(), o.extra.begin(), o.extra.end());
		vel.insert(vel.end(), o.vel.begin(), o.vel.end());
		general = general || o.general;
	}

	// class-4 fluid sites join the list; the list is kept in ascending site order
	std::sort(forced.begin(), forced.end());
	forced.erase(std::unique(forced.begin(), forced.end()), forced.end());
	list.insert(list.end(), forced.begin(), forced.end());
	std::sort(list.begin(), list.end());
	std::vector<int> extra_by_entry;
	if (!extra.empty())
	{
		extra_by_entry.assign(list.size(), 0);
		for (const auto &e : extra)
		{
			const size_t at = (size_t)(std::lower_bound(list.begin(), list.end(), e.first) - list.begin());
			extra_by_entry[at] += e.second;
		}
	}
	h->general = general || !forced.empty();

	cudaFree(h->bc_list); h->bc_list = nullptr;
	cudaFree(h->bc_extra); h->bc_extra = nullptr;
	cudaFree(h->vel_list); h->vel_list = nullptr;
	h->n_bc = (int)list.size();
	h->n_vel = (!reg && p.velocity_ramp_on) ? (int)vel.size() : 0;
	long long *forced_dev = nullptr;
	if (h->n_bc)
	{
		if (cudaMalloc(&h->bc_list, list.size() * sizeof(long long)) != cudaSuccess) FAIL(LUMA_B200_ENOMEM, "bc list");
		CK(cudaMemcpyAsync(h->bc_list, list.data(), list.size() * sizeof(long long), cudaMemcpyHostToDevice, h->s_main));
	}
	if (!extra_by_entry.empty())
	{
		if (cudaMalloc(&h->bc_extra, extra_by_entry.size() * sizeof(int)) != cudaSuccess) FAIL(LUMA_B200_ENOMEM, "bc extra");
		CK(cudaMemcpyAsync(h->bc_extra, extra_by_entry.data(), extra_by_entry.size() * sizeof(int), cudaMemcpyHostToDevice, h->s_main));
	}
	if (h->n_vel)
	{
		if (cudaMalloc(&h->vel_list, vel.size() * sizeof(long long)) != cudaSuccess) FAIL(LUMA_B200_ENOMEM, "velocity-site list");
		CK(cudaMemcpyAsync(h->vel_list, vel.data(), vel.size() * sizeof(long long), cudaMemcpyHostToDevice, h->s_main));
	}
	if (!forced.empty())
	{
		if (cudaMalloc(&forced_dev, forced.size() * sizeof(long long)) != cudaSuccess) FAIL(LUMA_B200_ENOMEM, "class-4 list");
		CK(cudaMemcpyAsync(forced_dev, forced.data(), forced.size() * sizeof(long long), cudaMemcpyHostToDevice, h->s_main));
	}
	GeomArgs g;
	g.types = h->types; g.bcdesc = h->bcdesc; g.cw = h->cw;
	g.P = h->P; g.M = p.M; g.K = p.K; g.wrap_x = h->ghost ? 0 : 1;
	g.p_begin = pb; g.p_end = pe;
	g.regularised = reg ? 1 : 0;
	LAT(h->Q, launch_cell_words<L>(g, h->s_main));
	h->st.kernel_launches++;
	if (!forced.empty())
	{
		int class_shift = 0;
		LAT(h->Q, class_shift = CW<L>::CLASS_SHIFT);
		launch_force_general(h->cw, forced_dev, (int)forced.size(), class_shift, h->s_main);
		h->st.kernel_launches++;
	}
	// Walls or bodies bounded along the fastest index (a fluid site whose z-neighbour -- y in 2-D -- is eSolid): k_step then completes the row-end sectors and
	// resolves bounce-back without splitting the warp (step_site / pull_warp).  Grids without them (periodic or open in z)
	// keep the plain sequence, which is ~1 % faster there (profiles/r02_probe_walls_after.txt).
	int *zflag = reinterpret_cast<int *>(h->momex_dev);      // scratch
	CK(cudaMemsetAsync(zflag, 0, sizeof(int), h->s_main));
	{
		// the two axis directions along the fastest-running index (z in 3-D, y in 2-D: K = 1)
		uint32_t zmask = 0;
		const int fast = (D == 3) ? 2 : 1;
		for (int v = 0; v < Q - 1; ++v)
		{
			int nz = 0;
			for (int d = 0; d < 3; ++d) nz += lat_c(Q, v, d) != 0;
			if (nz == 1 && lat_c(Q, v, fast) != 0) zmask |= 1u << v;
		}
		if (zmask) launch_any_bits(h->cw, (long long)pb * h->MK, (long long)(pe - pb) * h->MK, zmask, zflag, h->s_main);
	}
	int zwalls = 0;
	CK(cudaMemcpyAsync(&zwalls, zflag, sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
	const cudaError_t e1 = cudaGetLastError(), e2 = cudaStreamSynchronize(h->s_main);
	cudaFree(forced_dev);
	CK(e1); CK(e2);
	h->fill_holes = h->fill_env < 0 ? zwalls != 0 : h->fill_env != 0;
	return LUMA_B200_OK;
}
}  // extern "C++"

int luma_b200_upload(luma_b200_t *h, int32_t halo, const double *f_aos, const double *rho, const double *u_aos,
	const int32_t *lattyp, const LumaSiteBC *bc_sites, size_t n_bc,
	const double *ux_in, const double *uy_in, const double *uz_in)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!rho || !u_aos || !lattyp || halo < 0 || halo > 1) FAIL(LUMA_B200_EINVAL, "upload: null array or bad halo");
	if (n_bc && !bc_sites) FAIL(LUMA_B200_EINVAL, "upload: bc_sites");
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	// the state is replaced: steps accepted but not yet submitted are dropped, running ones are waited for
	h->t_enq = h->t; h->timing_open = false; h->stats_dirty = false; h->prof_used = 0;
	CK(cudaStreamSynchronize(h->s_comm));
	CK(cudaStreamSynchronize(h->s_main));
	h->stream_joined = true;
	// LUMA_B200_TRACE=1: wall time of the phases of this call on stderr (development aid)
	const bool trace = getenv("LUMA_B200_TRACE") != nullptr;
	auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t_begin = now();
	double t_copies = 0.0, t_desc = 0.0, t_geom = 0.0;
	const long long owned = (long long)p.x_count * h->MK;
	const long long host_off = (long long)halo * h->MK;       // first owned site in the host arrays
	const long long dev_off = (long long)h->ghost * h->MK;    // first owned site on the device
	const long long host_cells = owned + 2 * host_off;

	// populations: AoS chunks -> staging -> SoA lattice 0, then lattice 1 = lattice 0 (f.swap(fNew) keeps
	// never-updated sites identical in both, optimised.cpp:159 and init_grids.cpp:333)
	const long long chunk = std::max<long long>(h->MK, std::min<long long>(owned, (long long)(192u << 20) / (h->Q * 8)));
	int rc = ensure_staging(h, (size_t)chunk * h->Q * sizeof(double));
	if (rc) return rc;
	for (long long c0 = 0; f_aos && c0 < owned; c0 += chunk)
	{
		const long long n = std::min(chunk, owned - c0);
		CK(cudaMemcpyAsync(h->staging, f_aos + (host_off + c0) * h->Q, (size_t)n * h->Q * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
		LAT(h->Q, launch_aos_to_soa<L>((const double *)h->staging, h->f[0], h->stride, dev_off + c0, n, h->s_main));
		h->st.kernel_launches++;
	}
	for (long long c0 = 0; c0 < owned; c0 += chunk)
	{
		const long long n = std::min(chunk, owned - c0);
		CK(cudaMemcpyAsync(h->staging, u_aos + (host_off + c0) * h->D, (size_t)n * h->D * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
		launch_u_aos_to_soa((const double *)h->staging, h->u, h->stride, h->D, dev_off + c0, n, h->s_main);
		h->st.kernel_launches++;
	}
	CK(cudaMemcpyAsync(h->rho + dev_off, rho + host_off, (size_t)owned * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
	for (long long c0 = 0; c0 < owned; c0 += chunk * 4)
	{
		const long long n = std::min(chunk * 4, owned - c0);
		CK(cudaMemcpyAsync(h->staging, lattyp + host_off + c0, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->s_main));
		launch_types_from_i32((const int32_t *)h->staging, h->types + dev_off + c0, n, h->s_main);
		h->st.kernel_launches++;
	}
	if (!f_aos)
	{
		// f_aos == NULL: the host declares f = feq(rho, u) at every site, which is what LBM_initGrid leaves at t = 0
		// (src/GridObj_init_grids.cpp:310-333); the device evaluates the same expression bit for bit (k_feq_init uses
		// the step's own equilibrium_all) and 19 x 8 B per site stay off the PCIe bus
		LAT(h->Q, launch_feq_init<L>(h->rho, h->u, h->f[0], h->stride, dev_off, owned, h->C, h->s_main));
		h->st.kernel_launches++;
	}
	h->cur = 0;

	// inlet profiles (inc/GridObj.h:76-78)
	std::vector<double> uin((size_t)3 * p.M, 0.0);
	if (ux_in) memcpy(&uin[0], ux_in, sizeof(double) * p.M);
	if (uy_in) memcpy(&uin[p.M], uy_in, sizeof(double) * p.M);
	if (uz_in) memcpy(&uin[2 * (size_t)p.M], uz_in, sizeof(double) * p.M);
	CK(cudaMemcpyAsync(h->uin, uin.data(), uin.size() * sizeof(double), cudaMemcpyHostToDevice, h->s_main));

	if (trace) { cudaStreamSynchronize(h->s_main); t_copies = now(); }
	// wall descriptors of the boundary sites: sparse on the host, scattered into the dense device array
	std::vector<long long> dsite;
	std::vector<uint32_t> dval;
	dsite.reserve(n_bc); dval.reserve(n_bc);
	bool sorted = true;
	for (size_t b = 0; b < n_bc; ++b)
	{
		const LumaSiteBC &s = bc_sites[b];
		if (s.site < 0 || s.site >= host_cells) FAIL(LUMA_B200_EINVAL, "upload: bc site index out of range");
		const long long loc = s.site - host_off;
		if (loc < 0 || loc >= owned) continue;     // descriptor of a halo site: not ours
		if (s.edge_count < 0 || s.edge_count > 3 || s.normal_dir < 0 || s.normal_dir > 2) FAIL(LUMA_B200_EINVAL, "upload: bc descriptor");
		if (!s.edge_count) continue;
		if (!dsite.empty() && loc + dev_off <= dsite.back()) sorted = false;
		dsite.push_back(loc + dev_off);
		dval.push_back(cw_pack_bc(s.edge_count, s.normal_dir, s.normal[0], s.normal[1], s.normal[2]));
	}
	if (!sorted)
	{
		std::vector<size_t> order(dsite.size());
		for (size_t a = 0; a < order.size(); ++a) order[a] = a;
		std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return dsite[x] < dsite[y]; });
		std::vector<long long> s2(dsite.size()); std::vector<uint32_t> v2(dsite.size());
		for (size_t a = 0; a < order.size(); ++a) { s2[a] = dsite[order[a]]; v2[a] = dval[order[a]]; }
		dsite.swap(s2); dval.swap(v2);      // a site given twice keeps its last descriptor on the device (scatter order) and its first here
	}
	CK(cudaMemsetAsync(h->bcdesc, 0, (size_t)h->cells * sizeof(uint32_t), h->s_main));
	if (!dsite.empty())
	{
		const size_t need = dsite.size() * (sizeof(long long) + sizeof(uint32_t));
		rc = ensure_staging(h, need);
		if (rc) return rc;
		long long *ids_dev = (long long *)h->staging;
		uint32_t *val_dev = (uint32_t *)(ids_dev + dsite.size());
		CK(cudaMemcpyAsync(ids_dev, dsite.data(), dsite.size() * sizeof(long long), cudaMemcpyHostToDevice, h->s_main));
		CK(cudaMemcpyAsync(val_dev, dval.data(), dval.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->s_main));
		launch_scatter_u32(h->bcdesc, ids_dev, val_dev, (int)dsite.size(), h->s_main);
		h->st.kernel_launches++;
	}
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(h->s_main));
	t_desc = now();

	rc = finalize_geometry(h, [&](long long id, int, int, int) -> uint32_t
	{
		const auto it = std::lower_bound(dsite.begin(), dsite.end(), id);
		return (it != dsite.end() && *it == id) ? dval[(size_t)(it - dsite.begin())] : 0u;
	});
	if (rc) return rc;
	t_geom = now();
	if (h->ghost)
	{
		rc = exchange_populations(h, h->f[h->cur], h->s_main);
		if (rc) return rc;
	}
	// lattice 1 = lattice 0, ghost planes included (f.swap(fNew) keeps never-updated sites identical in both,
	// optimised.cpp:159 and init_grids.cpp:333)
	CK(cudaMemcpyAsync(h->f[1], h->f[0], (size_t)h->stride * h->Q * sizeof(double), cudaMemcpyDeviceToDevice, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	h->t = p.t; h->omega = p.omega; h->t_enq = h->t; h->omega_enq = h->omega;
	h->nu = (1.0 / h->omega - 0.5) * h->C.cs2;
	h->have_state = true; h->stepped = false;
	++h->geometry_epoch;
	if (trace)
		fprintf(stderr, "luma_b200_upload: %.1f ms = host->device copies + layout kernels %.1f, wall descriptors %.1f, geometry (eType scan, lists, cell words) %.1f, "
			"ghost exchange + second lattice %.1f\n", now() - t_begin, t_copies - t_begin, t_desc - t_copies, t_geom - t_desc, now() - t_geom);
	return halo_health(h);
}

int luma_b200_init_synthetic(luma_b200_t *h, const LumaSyntheticCase *c)
{
	if (!h || !c) return LUMA_B200_EINVAL;
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	h->t_enq = h->t; h->timing_open = false; h->stats_dirty = false; h->prof_used = 0;
	CK(cudaStreamSynchronize(h->s_comm));
	CK(cudaStreamSynchronize(h->s_main));
	h->stream_joined = true;
	for (int a = 0; a < 6; ++a)
	{
		const int t = c->wall_type[a];
		if (t != LUMA_E_SOLID && t != LUMA_E_FLUID && t != LUMA_E_VELOCITY && t != LUMA_E_PRESSURE && t != LUMA_E_SLIP && t != LUMA_E_EXTRAPOLATE_RIGHT)
			FAIL(LUMA_B200_EUNSUPPORTED, "wall type outside {eSolid,eFluid,eVelocity,ePressure,eSlip,eExtrapolateRight}");
		if (c->wall_cells[a] < 0) FAIL(LUMA_B200_EINVAL, "negative wall thickness");
	}
	std::vector<double> uin((size_t)3 * p.M);
	const double *prof[3] = { c->ux_in, c->uy_in, c->uz_in };
	for (int d = 0; d < 3; ++d) for (int j = 0; j < p.M; ++j)
		uin[(size_t)d * p.M + j] = prof[d] ? prof[d][j] : ((d < h->D) ? c->u_in[d] : 0.0);
	CK(cudaMemcpyAsync(h->uin, uin.data(), uin.size() * sizeof(double), cudaMemcpyHostToDevice, h->s_main));

	SynthArgs a;
	memset(&a, 0, sizeof(a));
	a.types = h->types; a.bcdesc = h->bcdesc; a.f0 = h->f[0]; a.f1 = h->f[1]; a.rho = h->rho; a.u = h->u;
	a.stride = h->stride; a.P = h->P; a.M = p.M; a.K = p.K; a.N = p.N;
	a.x_first = p.x_offset - h->ghost;
	for (int i = 0; i < 6; ++i) { a.wall_type[i] = c->wall_type[i]; a.wall_cells[i] = c->wall_cells[i]; a.box[i] = c->box[i]; }
	a.uin = h->uin;
	a.ramp0 = velocity_ramp_coef(p, 0.0);
	a.rhoin = p.rhoin;
	a.no_flow = c->no_flow; a.has_box = c->has_box;
	a.C = h->C;
	LAT(h->Q, launch_synthetic<L>(a, h->s_main));
	h->st.kernel_launches++;
	CK(cudaGetLastError());
	h->cur = 0;
	// the descriptor k_synthetic stores for a site, restated in cell indices (GridUtils::isWithinDomainWall)
	const int ghost = h->ghost, x_off = p.x_offset;
	int rc = finalize_geometry(h, [&, c](long long, int pl, int j, int k) -> uint32_t
	{
		const int gi = x_off + (pl - ghost);
		int ec = 0, nd = 0, n0 = 0, n1 = 0, n2 = 0;
		if (gi < c->wall_cells[0]) { nd = 0; n0 = 1; ++ec; }
		if (gi >= p.N - c->wall_cells[1]) { nd = 0; n0 = -1; ++ec; }
		if (j < c->wall_cells[2]) { nd = 1; n1 = 1; ++ec; }
		if (j >= p.M - c->wall_cells[3]) { nd = 1; n1 = -1; ++ec; }
		if (p.dims == 3)
		{
			if (k < c->wall_cells[4]) { nd = 2; n2 = 1; ++ec; }
			if (k >= p.K - c->wall_cells[5]) { nd = 2; n2 = -1; ++ec; }
		}
		return (ec > 0) ? cw_pack_bc(ec, nd, n0, n1, n2) : 0u;
	});
	if (rc) return rc;
	h->t = p.t; h->omega = p.omega; h->t_enq = h->t; h->omega_enq = h->omega;
	h->nu = (1.0 / h->omega - 0.5) * h->C.cs2;
	h->have_state = true; h->stepped = false;
	++h->geometry_epoch;
	return LUMA_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// The time step.  luma_b200_step() only ACCEPTS steps: it advances GridObj::t / omega / nu on the host and
// hands work to the GPU without ever waiting for it.  Work is submitted lazily so that
//   * the last step before the host looks at the fields is the one that stores rho,u (every other step keeps
//     them in registers), although the host calls LBM_multi_opt() once per step (src/main_lbm.cpp:441);
//   * launch-bound grids run as CUDA-graph batches even when the steps arrive one call at a time.
// Every entry point that reads state (download*, forces, stats, sync, flush) submits what is still deferred.
// ------------------------------------------------------------------------------------------------
static void fill_step_args(luma_b200_t *h, StepArgs &a)
{
	const LumaCaseParams &p = h->p;
	memset(&a, 0, sizeof(a));
	a.cw = h->cw; a.rho = h->rho; a.u = h->u; a.stride = h->stride;
	a.P = h->P; a.M = p.M; a.K = p.K; a.MK = (unsigned)h->MK;
	a.wrap_x = h->ghost ? 0 : 1;
	a.use_tma = h->use_tma;
	a.fill_holes = h->fill_holes ? 1 : 0;
	a.C = h->C;
	for (int v = 0; v < h->Q; ++v)
	{
		const int cx = lat_c(h->Q, v, 0), cy = lat_c(h->Q, v, 1), cz = lat_c(h->Q, v, 2);
		a.off_pull[v] = 8LL * ((long long)v * h->stride - ((long long)cx * h->MK + (long long)cy * p.K + cz));
	}
	a.bc_list = h->bc_list; a.bc_extra = h->bc_extra; a.n_bc = h->n_bc; a.uin = h->uin;
	a.rho_out = p.rho_out;
	a.types = h->types; a.bcdesc = h->bcdesc; a.general = h->general ? 1 : 0; a.regularised = p.regularised ? 1 : 0;
	a.velramp_on = p.velocity_ramp_on ? 1 : 0;
	a.tav = h->tav;
	// force_xyz = rho_init * gravity * refinement_ratio along L_GRAVITY_DIRECTION (init_grids.cpp:296-297)
	a.Fg = 0.0; a.hFg = 0.0;
	if (p.gravity_on) { a.Fg = p.rhoin * p.gravity * 1.0; a.hFg = 0.5 * a.Fg; }
	a.smag_coef = 2.0 * LUMA_SQRT2 * (p.csmag * p.csmag) * p.rhoin * h->C.cs2 * h->C.cs2;
}

// omega / nu of the step that takes the grid from t_now to t_now + 1:
// _LBM_updateReynolds (optimised.cpp:1313-1321), GridUnits::nud2nulbm (inc/GridUnits.h:128)
static void reynolds_step(const luma_b200_t *h, int t_now, double &omega, double &nu)
{
	const LumaCaseParams &p = h->p;
	if (!p.reynolds_ramp_on) return;
	const double newRe = p.re * reynolds_ramp_coef(p, (t_now + 1) * p.dt);
	nu = ((1.0 / newRe) * p.dt) / (p.dh * p.dh);
	omega = 1.0 / ((nu / h->C.cs2) + 0.5);
}

// per-step scalars of the step that takes the grid from time level t_now to t_now + 1
static void step_scalars(luma_b200_t *h, StepArgs &x, int t_now)
{
	const LumaCaseParams &p = h->p;
	double nu_unused = 0.0;
	reynolds_step(h, t_now, h->omega_enq, nu_unused);
	const double om = h->omega_enq;
	x.omega = om;
	x.tau = 1.0 / om;
	for (int k = 0; k < 4; ++k) x.lam[k] = (1 - 0.5 * om) * (h->C.w[k] / h->C.cs2);
	x.kbc_beta_m1 = 2.0 / om;
	x.kbc_inv_beta = 1.0 / x.kbc_beta_m1;
	x.ramp = velocity_ramp_coef(p, (t_now + 1) * p.dt);
	x.ramp_t = velocity_ramp_coef(p, t_now * p.dt);
	x.t_now = (double)t_now; x.t_next = (double)(t_now + 1);
}

static inline int force_code(const LumaCaseParams &p) { return p.gravity_on ? 1 + p.gravity_dir : 0; }
static inline int coll_code(const LumaCaseParams &p) { return p.kbc ? 2 : (p.bgksmag ? 1 : 0); }   // optimised.cpp:147-151

// One time step on the handle's two streams.
//   s_comm (highest priority): the list kernel k_bc (velocity / pressure / per-link sites), with slabs also the two
//           face planes and then the exchange of their outgoing populations;
//   s_main: k_step over the interior planes (all planes on a single rank).
// Both read lattice `fin` and write disjoint sites of `fout`, so the two streams run side by side and the exchange
// hides behind the interior kernel (no overlap exists in the reference: MpiManager.cpp:631 runs after :159).  Each
// stream waits for what the OTHER stream did in the previous step (ev_edge / ev_int) -- those kernels wrote sites
// this step reads and read sites this step overwrites.
static int enqueue_step_live(luma_b200_t *h, StepArgs &x, int coll, int force)
{
	const int owned = h->p.x_count;
	const bool side = h->ghost || h->n_bc > 0;
	if (side)
	{
		CK(cudaStreamWaitEvent(h->s_main, h->ev_edge, 0));
		CK(cudaStreamWaitEvent(h->s_comm, h->ev_int, 0));
		h->stream_joined = false;
	}
	if (!h->ghost)
	{
		x.p0 = 0; x.pstep = 1;
		if (side && h->profiling)
		{
			// per-kernel events on: the list kernel runs first on the main stream so that the events time k_step alone
			LAT(h->Q, launch_bc<L>(x, coll, force, h->s_main, &h->st.kernel_launches));
		}
		else if (side)
		{
			LAT(h->Q, launch_bc<L>(x, coll, force, h->s_comm, &h->st.kernel_launches));
			CK(cudaEventRecord(h->ev_edge, h->s_comm));
		}
		LAT(h->Q, main_kernel<L>(h, x, coll, force, h->P));
		if (side) CK(cudaEventRecord(h->ev_int, h->s_main));
		return LUMA_B200_OK;
	}
	const bool fused = h->p2p && h->fused;
	if (fused)
	{
		// the neighbours' copies of the lattice this step writes (the ranks step in lockstep: same lattice index)
		const int lo = (x.fout == h->f[0]) ? 0 : 1;
		for (int sd = 0; sd < 2; ++sd)
		{
			x.peer_f[sd] = h->peer[sd].f[lo];
			x.peer_stride[sd] = h->peer[sd].stride;
			x.peer_P[sd] = h->peer[sd].P;
		}
	}
	StepArgs e = x;
	e.p0 = 1; e.pstep = (owned > 1) ? owned - 1 : 1;
	const int nedge = (owned > 1) ? 2 : 1;
	StepArgs in = x;
	in.p0 = 2; in.pstep = 1;
	LAT(h->Q, launch_bc<L>(x, coll, force, h->s_comm, &h->st.kernel_launches));
	if (fused) LAT(h->Q, launch_step_faces<L>(e, coll, force, nedge, h->s_comm, &h->st.kernel_launches));
	else LAT(h->Q, launch_step<L>(e, coll, force, nedge, h->s_comm, &h->st.kernel_launches));
	CK(cudaEventRecord(h->ev_edge, h->s_comm));
	if (h->profiling) CK(cudaStreamWaitEvent(h->s_main, h->ev_edge, 0));      // events on: the interior kernel is timed alone
	const int rc = exchange_populations(h, x.fout, h->s_comm, fused);
	if (rc) return rc;
	CK(cudaEventRecord(h->ev_comm, h->s_comm));
	LAT(h->Q, main_kernel<L>(h, in, coll, force, owned - 2));
	CK(cudaEventRecord(h->ev_int, h->s_main));
	return LUMA_B200_OK;
}

// the same step as a self-contained fork/join on the capturing stream: the body of the CUDA graphs.  Slabs: only with the
// device-initiated exchange (its launches take no per-step argument: the exchange number lives on the device).
static int enqueue_step_captured(luma_b200_t *h, StepArgs &x, int coll, int force)
{
	if (!h->ghost)
	{
		x.p0 = 0; x.pstep = 1;
		const bool side = h->n_bc > 0;
		if (side)
		{
			CK(cudaEventRecord(h->ev_fork, h->s_main));
			CK(cudaStreamWaitEvent(h->s_comm, h->ev_fork, 0));
			LAT(h->Q, launch_bc<L>(x, coll, force, h->s_comm, &h->st.kernel_launches));
			CK(cudaEventRecord(h->ev_join, h->s_comm));
		}
		LAT(h->Q, launch_step<L>(x, coll, force, h->P, h->s_main, &h->st.kernel_launches));
		if (side) CK(cudaStreamWaitEvent(h->s_main, h->ev_join, 0));
		return LUMA_B200_OK;
	}
	const int owned = h->p.x_count;
	const bool fused = h->fused;
	if (fused)
	{
		const int lo = (x.fout == h->f[0]) ? 0 : 1;
		for (int sd = 0; sd < 2; ++sd)
		{
			x.peer_f[sd] = h->peer[sd].f[lo];
			x.peer_stride[sd] = h->peer[sd].stride;
			x.peer_P[sd] = h->peer[sd].P;
		}
	}
	StepArgs e = x;
	e.p0 = 1; e.pstep = (owned > 1) ? owned - 1 : 1;
	const int nedge = (owned > 1) ? 2 : 1;
	StepArgs in = x;
	in.p0 = 2; in.pstep = 1;
	CK(cudaEventRecord(h->ev_fork, h->s_main));
	CK(cudaStreamWaitEvent(h->s_comm, h->ev_fork, 0));
	LAT(h->Q, launch_bc<L>(x, coll, force, h->s_comm, &h->st.kernel_launches));
	if (fused) LAT(h->Q, launch_step_faces<L>(e, coll, force, nedge, h->s_comm, &h->st.kernel_launches));
	else LAT(h->Q, launch_step<L>(e, coll, force, nedge, h->s_comm, &h->st.kernel_launches));
	const int rc = exchange_populations(h, x.fout, h->s_comm, fused);
	if (rc) return rc;
	CK(cudaEventRecord(h->ev_join, h->s_comm));
	LAT(h->Q, launch_step<L>(in, coll, force, owned - 2, h->s_main, &h->st.kernel_launches));
	CK(cudaStreamWaitEvent(h->s_main, h->ev_join, 0));
	return LUMA_B200_OK;
}

// s_main waits for everything issued on s_comm (so that work queued on s_main afterwards sees a complete step)
static int join_streams(luma_b200_t *h)
{
	if (h->stream_joined) return LUMA_B200_OK;
	CK(cudaStreamWaitEvent(h->s_main, h->ev_edge, 0));
	if (h->ghost) CK(cudaStreamWaitEvent(h->s_main, h->ev_comm, 0));
	h->stream_joined = true;
	return LUMA_B200_OK;
}

// device time of the steps between two read points (ev_t0 .. ev_t1) -> LumaStats, without blocking unless asked to
static int collect_stats(luma_b200_t *h, bool wait)
{
	if (!h->stats_dirty) return LUMA_B200_OK;
	if (!wait && cudaEventQuery(h->ev_t1) != cudaSuccess) { cudaGetLastError(); return LUMA_B200_OK; }
	CK(cudaEventSynchronize(h->ev_t1));
	float ms = 0.f;
	CK(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
	for (size_t e = 0; e + 1 < h->prof_used; e += 2)
	{
		float kms = 0.f;
		CK(cudaEventElapsedTime(&kms, h->prof_ev[e], h->prof_ev[e + 1]));
		h->st.step_kernel_ms += kms;
	}
	h->prof_used = 0;
	h->stats_dirty = false;
	const int n = h->window_steps > 0 ? h->window_steps : 1;
	h->st.ms_last_call = ms;
	h->st.ms_per_step = ms / n;
	h->st.mlups_last_call = (double)h->st.cells * n / ((double)ms * 1e3);
	return LUMA_B200_OK;
}

// submit deferred steps.  final = false: keep the last accepted step back (it may have to store rho,u), and on
// launch-bound grids wait for a full CUDA-graph batch; final = true: submit everything, the last step stores rho,u.
static int drain(luma_b200_t *h, bool final)
{
	int pending = h->t - h->t_enq;
	if (pending <= 0) return LUMA_B200_OK;
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	const int force = force_code(p), coll = coll_code(p);
	const int GS = h->graph_steps;
	// Launch-bound grids (BASELINE configs[0], 256^2: ~3 us of work per step): batches of GS steps are captured once
	// into a CUDA graph -- same kernels, same arguments, the two streams become graph branches -- and replayed with
	// one launch each.  Only while every per-step scalar is constant (ramps finished, no time averages, no profiling
	// events), on a single rank, and never for a step that stores rho,u.
	// Slabs qualify with the device-initiated exchange (LUMA_B200_GRAPH_SLABS=0 keeps them on live launches).
	const bool graph_capable = GS >= 2 && (!h->ghost || (h->p2p && h->graph_slabs)) && !h->profiling && !h->tav;
	auto graph_ready = [&](int t_now) -> bool
	{
		if (!graph_capable) return false;
		if (velocity_ramp_coef(p, t_now * p.dt) != 1.0 || velocity_ramp_coef(p, (t_now + 1) * p.dt) != 1.0) return false;
		if (p.reynolds_ramp_on && !((t_now + 1) * p.dt > p.reynolds_ramp)) return false;
		return true;
	};
	if (!final && (pending <= 1 || (graph_ready(h->t_enq) && pending - 1 < GS))) return LUMA_B200_OK;

	StepArgs a;
	fill_step_args(h, a);
	if (!h->timing_open)
	{
		// a new timing window opens: the previous one is read now if it has completed, dropped otherwise
		int rc = collect_stats(h, h->profiling);
		if (rc) return rc;
		h->stats_dirty = false; h->prof_used = 0;
		CK(cudaEventRecord(h->ev_t0, h->s_main));
		h->timing_open = true; h->timed_steps = 0;
	}
	CK(cudaEventRecord(h->ev_int, h->s_main));      // whatever ran on s_main since the last step (uploads, snapshots ...)
	const int keep = final ? 0 : 1;
	while (pending > keep)
	{
		if (pending - 1 >= GS && graph_ready(h->t_enq))
		{
			StepArgs b = a;
			step_scalars(h, b, h->t_enq);
			GraphSlot &gs = h->graphs[h->cur];
			if (gs.exec && (gs.omega != b.omega || gs.n_bc != h->n_bc || gs.epoch != h->geometry_epoch))
			{
				cudaGraphExecDestroy(gs.exec); gs.exec = nullptr;
			}
			int rc = join_streams(h);
			if (rc) return rc;
			if (!gs.exec)
			{
				const int64_t before = h->st.kernel_launches;
				cudaGraph_t graph = nullptr;
				CK(cudaStreamBeginCapture(h->s_main, cudaStreamCaptureModeRelaxed));
				for (int i = 0; i < GS && rc == LUMA_B200_OK; ++i)
				{
					b.fin = h->f[h->cur ^ (i & 1)]; b.fout = h->f[h->cur ^ (i & 1) ^ 1];
					b.write_macro = 0;
					rc = enqueue_step_captured(h, b, coll, force);
				}
				const cudaError_t ce = cudaStreamEndCapture(h->s_main, &graph);
				gs.nodes = h->st.kernel_launches - before;
				h->st.kernel_launches = before;
				if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
				CK(ce);
				const cudaError_t ie = cudaGraphInstantiate(&gs.exec, graph, 0);
				cudaGraphDestroy(graph);
				CK(ie);
				gs.omega = b.omega; gs.n_bc = h->n_bc; gs.epoch = h->geometry_epoch;
			}
			CK(cudaGraphLaunch(gs.exec, h->s_main));
			CK(cudaEventRecord(h->ev_int, h->s_main));
			h->st.kernel_launches += gs.nodes;
			h->st.graph_launches++;
			h->t_enq += GS;       // GS is even: h->cur is unchanged
			h->timed_steps += GS;
			pending -= GS;
			continue;
		}
		if (!final && graph_ready(h->t_enq)) break;      // fewer than a batch left: wait for more steps or for a read point
		step_scalars(h, a, h->t_enq);
		a.fin = h->f[h->cur]; a.fout = h->f[h->cur ^ 1];
		a.write_macro = (final && pending == 1) ? 1 : 0;
		{
			const int rc = enqueue_step_live(h, a, coll, force);
			if (rc) return rc;
		}
		if (a.write_macro && h->n_vel)
		{
			// stored u of the forced-equilibrium inlet sites as the reference leaves it after this step
			int rc = join_streams(h);
			if (rc) return rc;
			VelSrcArgs vs;
			vs.list = h->vel_list; vs.n = h->n_vel; vs.types = h->types; vs.bcdesc = h->bcdesc; vs.u = h->u; vs.stride = h->stride;
			vs.uin = h->uin; vs.ramp_t = a.ramp_t;
			vs.P = h->P; vs.M = p.M; vs.K = p.K; vs.N = p.N; vs.wrap_x = a.wrap_x; vs.x_first = p.x_offset - h->ghost;
			LAT(h->Q, launch_velsrc<L>(vs, h->s_main, &h->st.kernel_launches));
		}
		h->cur ^= 1;
		++h->t_enq;
		++h->timed_steps;
		--pending;
		h->stepped = true;
		if ((h->timed_steps & 63) == 0) { const int rc = halo_health(h); if (rc) return rc; }
	}
	CK(cudaGetLastError());
	return LUMA_B200_OK;
}

// read point: everything accepted so far is submitted, the streams are joined into s_main and the timing window is
// closed (ev_t1); nothing here waits for the GPU
static int flush_steps(luma_b200_t *h)
{
	int rc = drain(h, true);
	if (rc) return rc;
	rc = join_streams(h);
	if (rc) return rc;
	if (h->timing_open)
	{
		CK(cudaEventRecord(h->ev_t1, h->s_main));
		h->timing_open = false;
		h->stats_dirty = true;
		h->window_steps = h->timed_steps;
	}
	return LUMA_B200_OK;
}

int luma_b200_set_profiling(luma_b200_t *h, int32_t on)
{
	if (!h) return LUMA_B200_EINVAL;
	if (h->have_state)
	{
		int rc = flush_steps(h);
		if (rc == LUMA_B200_OK) rc = collect_stats(h, true);
		if (rc) return rc;
	}
	h->profiling = on != 0;
	h->st.step_kernel_launches = 0; h->st.step_kernel_ms = 0.0; h->st.step_kernel_cells = 0;
	return LUMA_B200_OK;
}

int luma_b200_step(luma_b200_t *h, int32_t nsteps)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "step before upload/init_synthetic");
	if (nsteps < 0) FAIL(LUMA_B200_EINVAL, "nsteps < 0");
	if (nsteps == 0) return LUMA_B200_OK;
	int rc = halo_health(h);
	if (rc) return rc;
	// GridObj::t, ::omega, ::nu as the reference leaves them after these steps (optimised.cpp:39-42, :170)
	h->t += nsteps;
	reynolds_step(h, h->t - 1, h->omega, h->nu);
	h->st.steps += nsteps;
	int plus[9];
	h->st.halo_bytes_per_step = h->ghost ? 2LL * edge_pops(h->Q, +1, plus) * h->MK * (long long)sizeof(double) : 0;
	return drain(h, false);
}

int luma_b200_flush(luma_b200_t *h)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!h->have_state) return LUMA_B200_OK;
	CK(cudaSetDevice(h->p.device));
	return flush_steps(h);
}

int luma_b200_download(luma_b200_t *h, int32_t halo, unsigned what, double *f_aos, double *rho, double *u_aos)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "download before upload/init_synthetic");
	if (halo < 0 || halo > 1) FAIL(LUMA_B200_EINVAL, "download: halo");
	if (((what & LUMA_B200_F) && !f_aos) || ((what & LUMA_B200_RHO) && !rho) || ((what & LUMA_B200_U) && !u_aos))
		FAIL(LUMA_B200_EINVAL, "download: null array");
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	int rc = flush_steps(h);
	if (rc) return rc;
	const long long owned = (long long)p.x_count * h->MK;
	const long long host_off = (long long)halo * h->MK, dev_off = (long long)h->ghost * h->MK;
	const long long chunk = std::max<long long>(h->MK, std::min<long long>(owned, (long long)(192u << 20) / (h->Q * 8)));
	rc = ensure_staging(h, (size_t)chunk * h->Q * sizeof(double));
	if (rc) return rc;
	if (what & LUMA_B200_F)
		for (long long c0 = 0; c0 < owned; c0 += chunk)
		{
			const long long n = std::min(chunk, owned - c0);
			LAT(h->Q, launch_soa_to_aos<L>(h->f[h->cur], (double *)h->staging, h->stride, dev_off + c0, n, h->s_main));
			h->st.kernel_launches++;
			CK(cudaMemcpyAsync(f_aos + (host_off + c0) * h->Q, h->staging, (size_t)n * h->Q * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
		}
	if (what & LUMA_B200_U)
		for (long long c0 = 0; c0 < owned; c0 += chunk)
		{
			const long long n = std::min(chunk, owned - c0);
			launch_u_soa_to_aos(h->u, (double *)h->staging, h->stride, h->D, dev_off + c0, n, h->s_main);
			h->st.kernel_launches++;
			CK(cudaMemcpyAsync(u_aos + (host_off + c0) * h->D, h->staging, (size_t)n * h->D * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
		}
	if (what & LUMA_B200_RHO)
		CK(cudaMemcpyAsync(rho + host_off, h->rho + dev_off, (size_t)owned * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(h->s_main));
	return halo_health(h);
}

// Asynchronous variant for hosts that write their output while the next steps run: the requested fields
// are snapshotted on the device in the host layout (stream-ordered after the steps so far, a few hundred
// microseconds), then copied to the host on a separate stream while luma_b200_step keeps the GPU busy.
int luma_b200_download_async(luma_b200_t *h, int32_t halo, unsigned what, double *f_aos, double *rho, double *u_aos)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "download before upload/init_synthetic");
	if (halo < 0 || halo > 1) FAIL(LUMA_B200_EINVAL, "download: halo");
	if (((what & LUMA_B200_F) && !f_aos) || ((what & LUMA_B200_RHO) && !rho) || ((what & LUMA_B200_U) && !u_aos))
		FAIL(LUMA_B200_EINVAL, "download: null array");
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	{
		const int rc = flush_steps(h);
		if (rc) return rc;
	}
	const long long owned = (long long)p.x_count * h->MK;
	const long long host_off = (long long)halo * h->MK, dev_off = (long long)h->ghost * h->MK;
	size_t need = 0;
	if (what & LUMA_B200_RHO) need += (size_t)owned;
	if (what & LUMA_B200_U) need += (size_t)owned * h->D;
	if (what & LUMA_B200_F) need += (size_t)owned * h->Q;
	need *= sizeof(double);
	if (h->copy_pending) CK(cudaStreamWaitEvent(h->s_main, h->ev_copied, 0));      // the previous copy still reads the snapshot
	if (h->snap_bytes < need)
	{
		if (h->copy_pending) CK(cudaEventSynchronize(h->ev_copied));
		cudaFree(h->snap); h->snap = nullptr; h->snap_bytes = 0;
		if (cudaMalloc(&h->snap, need) != cudaSuccess) { cudaGetLastError(); FAIL(LUMA_B200_ENOMEM, "snapshot buffer"); }
		h->snap_bytes = need;
	}
	double *s_rho = h->snap, *s_u = s_rho + ((what & LUMA_B200_RHO) ? owned : 0), *s_f = s_u + ((what & LUMA_B200_U) ? owned * h->D : 0);
	if (what & LUMA_B200_RHO)
		CK(cudaMemcpyAsync(s_rho, h->rho + dev_off, (size_t)owned * sizeof(double), cudaMemcpyDeviceToDevice, h->s_main));
	if (what & LUMA_B200_U)
	{
		launch_u_soa_to_aos(h->u, s_u, h->stride, h->D, dev_off, owned, h->s_main);
		h->st.kernel_launches++;
	}
	if (what & LUMA_B200_F)
	{
		LAT(h->Q, launch_soa_to_aos<L>(h->f[h->cur], s_f, h->stride, dev_off, owned, h->s_main));
		h->st.kernel_launches++;
	}
	CK(cudaGetLastError());
	CK(cudaEventRecord(h->ev_snap, h->s_main));
	CK(cudaStreamWaitEvent(h->s_copy, h->ev_snap, 0));
	if (what & LUMA_B200_RHO) CK(cudaMemcpyAsync(rho + host_off, s_rho, (size_t)owned * sizeof(double), cudaMemcpyDeviceToHost, h->s_copy));
	if (what & LUMA_B200_U) CK(cudaMemcpyAsync(u_aos + host_off * h->D, s_u, (size_t)owned * h->D * sizeof(double), cudaMemcpyDeviceToHost, h->s_copy));
	if (what & LUMA_B200_F) CK(cudaMemcpyAsync(f_aos + host_off * h->Q, s_f, (size_t)owned * h->Q * sizeof(double), cudaMemcpyDeviceToHost, h->s_copy));
	CK(cudaEventRecord(h->ev_copied, h->s_copy));
	h->copy_pending = true;
	return LUMA_B200_OK;
}

int luma_b200_download_wait(luma_b200_t *h)
{
	if (!h) return LUMA_B200_EINVAL;
	CK(cudaSetDevice(h->p.device));
	if (h->copy_pending) CK(cudaEventSynchronize(h->ev_copied));
	h->copy_pending = false;
	return halo_health(h);
}

// rho_timeav [cells], ui_timeav [cells*D], uiuj_timeav [cells*(3D-3)] in the reference's AoS layout (inc/GridObj.h:93-95)
static int transfer_timeav(luma_b200_t *h, int32_t halo, double *rho_tav, double *ui_tav, double *uiuj_tav, bool to_host)
{
	if (!h->tav) FAIL(LUMA_B200_ESTATE, "the handle was created without time_averaged");
	if (halo < 0 || halo > 1) FAIL(LUMA_B200_EINVAL, "timeav: halo");
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	int rc = flush_steps(h);
	if (rc) return rc;
	const long long owned = (long long)p.x_count * h->MK;
	const long long host_off = (long long)halo * h->MK, dev_off = (long long)h->ghost * h->MK;
	const long long chunk = std::max<long long>(h->MK, std::min<long long>(owned, (long long)(192u << 20) / (h->Q * 8)));
	rc = ensure_staging(h, (size_t)chunk * h->Q * sizeof(double));
	if (rc) return rc;
	if (rho_tav)
	{
		if (to_host) CK(cudaMemcpyAsync(rho_tav + host_off, h->tav + dev_off, (size_t)owned * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
		else CK(cudaMemcpyAsync(h->tav + dev_off, rho_tav + host_off, (size_t)owned * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
	}
	double *host[2] = { ui_tav, uiuj_tav };
	const int ncomp[2] = { h->D, 3 * h->D - 3 }, slot[2] = { 1, 1 + h->D };
	for (int a = 0; a < 2; ++a)
	{
		if (!host[a]) continue;
		double *soa = h->tav + (long long)slot[a] * h->stride;
		for (long long c0 = 0; c0 < owned; c0 += chunk)
		{
			const long long n = std::min(chunk, owned - c0);
			double *hp = host[a] + (host_off + c0) * ncomp[a];
			if (to_host)
			{
				launch_u_soa_to_aos(soa, (double *)h->staging, h->stride, ncomp[a], dev_off + c0, n, h->s_main);
				CK(cudaMemcpyAsync(hp, h->staging, (size_t)n * ncomp[a] * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
			}
			else
			{
				CK(cudaMemcpyAsync(h->staging, hp, (size_t)n * ncomp[a] * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
				launch_u_aos_to_soa((const double *)h->staging, soa, h->stride, ncomp[a], dev_off + c0, n, h->s_main);
			}
			h->st.kernel_launches++;
		}
	}
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(h->s_main));
	return LUMA_B200_OK;
}

int luma_b200_download_timeav(luma_b200_t *h, int32_t halo, double *rho_timeav, double *ui_timeav, double *uiuj_timeav)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "download before upload/init_synthetic");
	return transfer_timeav(h, halo, rho_timeav, ui_timeav, uiuj_timeav, true);
}

int luma_b200_upload_timeav(luma_b200_t *h, int32_t halo, const double *rho_timeav, const double *ui_timeav, const double *uiuj_timeav)
{
	if (!h) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "upload_timeav before upload/init_synthetic");
	return transfer_timeav(h, halo, const_cast<double *>(rho_timeav), const_cast<double *>(ui_timeav), const_cast<double *>(uiuj_timeav), false);
}

// ------------------------------------------------------------------------------------------------
// Binary restart.  GridObj::io_restart (src/GridObj_ops_io.cpp:406-640) writes / reads the level-0 state as
// ASCII, one line per site (indices, Q populations, rho, u, time averages); this is the same content -- t, f, rho, u and,
// when the handle has them, rho_timeav / ui_timeav / uiuj_timeav of this rank's owned planes, in the reference's array
// layouts -- as raw little-endian doubles behind a small header, streamed through the staging buffer in chunks.
// ------------------------------------------------------------------------------------------------
struct RestartHeader
{
	char magic[8];            // "LUMAB2R1"
	int32_t dims, num_vels, N, M, K, x_offset, x_count, t;
	int32_t has_timeav, reserved;
	double omega, nu;
};

static int restart_stream(luma_b200_t *h, FILE *fh, bool write, double *soa_or_field, int ncomp, bool soa)
{
	// one field of `ncomp` doubles per site (ncomp = 1: plain copy; > 1: device SoA <-> file AoS), owned planes only
	const long long owned = (long long)h->p.x_count * h->MK, dev_off = (long long)h->ghost * h->MK;
	const long long chunk = std::max<long long>(h->MK, std::min<long long>(owned, (long long)(64u << 20) / (ncomp * 8)));
	if (soa)
	{
		const int rc = ensure_staging(h, (size_t)chunk * ncomp * sizeof(double));      // layout conversion goes through the staging buffer
		if (rc) return rc;
	}
	std::vector<double> host((size_t)chunk * ncomp);
	for (long long c0 = 0; c0 < owned; c0 += chunk)
	{
		const long long n = std::min(chunk, owned - c0);
		const size_t bytes = (size_t)n * ncomp * sizeof(double);
		if (write)
		{
			if (!soa) CK(cudaMemcpyAsync(host.data(), soa_or_field + dev_off + c0, bytes, cudaMemcpyDeviceToHost, h->s_main));
			else
			{
				if (ncomp == h->Q) LAT(h->Q, launch_soa_to_aos<L>(soa_or_field, (double *)h->staging, h->stride, dev_off + c0, n, h->s_main));
				else launch_u_soa_to_aos(soa_or_field, (double *)h->staging, h->stride, ncomp, dev_off + c0, n, h->s_main);
				CK(cudaMemcpyAsync(host.data(), h->staging, bytes, cudaMemcpyDeviceToHost, h->s_main));
			}
			CK(cudaStreamSynchronize(h->s_main));
			if (fwrite(host.data(), 1, bytes, fh) != bytes) FAIL(LUMA_B200_EINVAL, "restart: short write");
		}
		else
		{
			if (fread(host.data(), 1, bytes, fh) != bytes) FAIL(LUMA_B200_EINVAL, "restart: file too short");
			if (!soa) CK(cudaMemcpyAsync(soa_or_field + dev_off + c0, host.data(), bytes, cudaMemcpyHostToDevice, h->s_main));
			else
			{
				CK(cudaMemcpyAsync(h->staging, host.data(), bytes, cudaMemcpyHostToDevice, h->s_main));
				if (ncomp == h->Q) LAT(h->Q, launch_aos_to_soa<L>((const double *)h->staging, soa_or_field, h->stride, dev_off + c0, n, h->s_main));
				else launch_u_aos_to_soa((const double *)h->staging, soa_or_field, h->stride, ncomp, dev_off + c0, n, h->s_main);
			}
			CK(cudaStreamSynchronize(h->s_main));
		}
		h->st.kernel_launches += soa ? 1 : 0;
	}
	return LUMA_B200_OK;
}

static int restart_fields(luma_b200_t *h, FILE *fh, bool write)
{
	int rc = restart_stream(h, fh, write, h->rho, 1, false);
	if (rc == LUMA_B200_OK) rc = restart_stream(h, fh, write, h->u, h->D, true);
	if (rc == LUMA_B200_OK) rc = restart_stream(h, fh, write, h->f[h->cur], h->Q, true);
	if (rc == LUMA_B200_OK && h->tav)
	{
		rc = restart_stream(h, fh, write, h->tav, 1, false);
		if (rc == LUMA_B200_OK) rc = restart_stream(h, fh, write, h->tav + (long long)1 * h->stride, h->D, true);
		if (rc == LUMA_B200_OK) rc = restart_stream(h, fh, write, h->tav + (long long)(1 + h->D) * h->stride, 3 * h->D - 3, true);
	}
	return rc;
}

int luma_b200_restart_write(luma_b200_t *h, const char *path)
{
	if (!h || !path) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "restart_write before upload/init_synthetic");
	CK(cudaSetDevice(h->p.device));
	int rc = flush_steps(h);
	if (rc) return rc;
	FILE *fh = fopen(path, "wb");
	if (!fh) FAIL(LUMA_B200_EINVAL, std::string("restart_write: cannot open ") + path);
	RestartHeader hd;
	memset(&hd, 0, sizeof(hd));
	memcpy(hd.magic, "LUMAB2R1", 8);
	hd.dims = h->D; hd.num_vels = h->Q; hd.N = h->p.N; hd.M = h->p.M; hd.K = h->p.K;
	hd.x_offset = h->p.x_offset; hd.x_count = h->p.x_count; hd.t = h->t; hd.has_timeav = h->tav ? 1 : 0;
	hd.omega = h->omega; hd.nu = h->nu;
	rc = fwrite(&hd, sizeof(hd), 1, fh) == 1 ? LUMA_B200_OK : LUMA_B200_EINVAL;
	if (rc == LUMA_B200_OK) rc = restart_fields(h, fh, true);
	if (fclose(fh) != 0 && rc == LUMA_B200_OK) { h->err = "restart_write: close failed"; rc = LUMA_B200_EINVAL; }
	return rc;
}

int luma_b200_restart_read(luma_b200_t *h, const char *path)
{
	if (!h || !path) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "restart_read needs the geometry first (upload or init_synthetic)");
	CK(cudaSetDevice(h->p.device));
	// the state is replaced: steps accepted but not yet submitted are dropped, running ones are waited for
	h->t_enq = h->t; h->timing_open = false; h->stats_dirty = false; h->prof_used = 0;
	CK(cudaStreamSynchronize(h->s_comm));
	CK(cudaStreamSynchronize(h->s_main));
	h->stream_joined = true;
	FILE *fh = fopen(path, "rb");
	if (!fh) FAIL(LUMA_B200_EINVAL, std::string("restart_read: cannot open ") + path);
	RestartHeader hd;
	int rc = LUMA_B200_OK;
	if (fread(&hd, sizeof(hd), 1, fh) != 1 || memcmp(hd.magic, "LUMAB2R1", 8) != 0) { h->err = "restart_read: not a luma_b200 restart file"; rc = LUMA_B200_EINVAL; }
	else if (hd.dims != h->D || hd.num_vels != h->Q || hd.N != h->p.N || hd.M != h->p.M || hd.K != h->p.K ||
		hd.x_offset != h->p.x_offset || hd.x_count != h->p.x_count || (hd.has_timeav != 0) != (h->tav != nullptr))
	{ h->err = "restart_read: the file was written for another grid, slab or build (time averages)"; rc = LUMA_B200_EINVAL; }
	if (rc == LUMA_B200_OK) rc = restart_fields(h, fh, false);
	fclose(fh);
	if (rc) return rc;
	if (h->ghost)
	{
		rc = exchange_populations(h, h->f[h->cur], h->s_main);
		if (rc) return rc;
	}
	CK(cudaMemcpyAsync(h->f[h->cur ^ 1], h->f[h->cur], (size_t)h->stride * h->Q * sizeof(double), cudaMemcpyDeviceToDevice, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	h->t = hd.t; h->t_enq = hd.t; h->omega = hd.omega; h->omega_enq = hd.omega; h->nu = hd.nu;
	h->stepped = false;
	return halo_health(h);
}

int luma_b200_download_lattyp(luma_b200_t *h, int32_t halo, int32_t *lattyp)
{
	if (!h || !lattyp) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "download before upload/init_synthetic");
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	const long long owned = (long long)p.x_count * h->MK;
	std::vector<uint8_t> t((size_t)owned);
	CK(cudaMemcpyAsync(t.data(), h->types + (long long)h->ghost * h->MK, (size_t)owned, cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaStreamSynchronize(h->s_main));
	int32_t *dst = lattyp + (long long)halo * h->MK;
	for (long long i = 0; i < owned; ++i) dst[i] = (int32_t)t[(size_t)i];
	return LUMA_B200_OK;
}

int luma_b200_get_time(luma_b200_t *h, int32_t *t, double *omega, double *nu)
{
	if (!h) return LUMA_B200_EINVAL;
	if (t) *t = h->t;
	if (omega) *omega = h->omega;
	if (nu) *nu = h->nu;
	return LUMA_B200_OK;
}

int luma_b200_forces(luma_b200_t *h, double F[3])
{
	if (!h || !F) return LUMA_B200_EINVAL;
	if (!h->have_state) FAIL(LUMA_B200_ESTATE, "forces before upload/init_synthetic");
	const LumaCaseParams &p = h->p;
	CK(cudaSetDevice(p.device));
	int rc = flush_steps(h);
	if (rc) return rc;
	if (!h->stepped) FAIL(LUMA_B200_ESTATE, "forces need at least one step (they use the pre-stream populations of the last step)");
	// Every link is summed by the rank that owns its FLUID end: the populations read are those of this rank's own
	// planes of the previous lattice, which nobody else writes (a ring neighbour that is already one step ahead
	// stores into the GHOST planes of that lattice); the solid end may lie in a ghost plane (eType is static).
	const double *prev = h->f[h->cur ^ 1];
	int nb;
	LAT(h->Q, nb = launch_momex<L>(prev, h->types, h->stride, h->P, p.M, p.K, h->ghost, h->P - h->ghost, p.x_offset - h->ghost, p.N, h->momex_dev, 4096, h->s_main));
	h->st.kernel_launches++;
	std::vector<double> part((size_t)3 * nb);
	CK(cudaMemcpyAsync(part.data(), h->momex_dev, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->s_main));
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(h->s_main));
	F[0] = F[1] = F[2] = 0.0;
	for (int b = 0; b < nb; ++b) { F[0] += part[3 * b]; F[1] += part[3 * b + 1]; F[2] += part[3 * b + 2]; }
	return halo_health(h);
}

int luma_b200_stats(luma_b200_t *h, LumaStats *s)
{
	if (!h || !s) return LUMA_B200_EINVAL;
	if (h->have_state)
	{
		// a read point: what is still deferred is submitted, then the device time of the steps since the previous
		// read point is collected (this is the one place that waits for them)
		CK(cudaSetDevice(h->p.device));
		int rc = flush_steps(h);
		if (rc == LUMA_B200_OK) rc = collect_stats(h, true);
		if (rc == LUMA_B200_OK) rc = halo_health(h);
		if (rc) return rc;
	}
	*s = h->st;
	return LUMA_B200_OK;
}

int luma_b200_halo_plan(const LumaCaseParams *p, LumaHaloMsg *msgs, int32_t capacity, int32_t *count)
{
	if (!p || !count || p->struct_size != sizeof(LumaCaseParams)) return LUMA_B200_EINVAL;
	if (!((p->dims == 3 && (p->num_vels == 19 || p->num_vels == 27)) || (p->dims == 2 && p->num_vels == 9))) return LUMA_B200_EINVAL;
	if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks || p->x_count < 1) return LUMA_B200_EINVAL;
	std::vector<LumaHaloMsg> plan;
	const int n = build_halo_plan(*p, plan);
	*count = n;
	if (!msgs) return LUMA_B200_OK;
	if (capacity < n) return LUMA_B200_EINVAL;
	for (int i = 0; i < n; ++i) msgs[i] = plan[(size_t)i];
	return LUMA_B200_OK;
}

int luma_b200_selftest_div_const(int32_t device, int64_t n, uint64_t seed, int64_t *mismatches)
{
	if (!mismatches || n < 0) return LUMA_B200_EINVAL;
	if (cudaSetDevice(device) != cudaSuccess) return LUMA_B200_ECUDA;
	LbmConst C;
	make_constants(C, 19);
	unsigned long long *d = nullptr, hcount = 0;
	if (cudaMalloc(&d, sizeof(*d)) != cudaSuccess) return LUMA_B200_ENOMEM;
	cudaMemset(d, 0, sizeof(*d));
	launch_selftest_div(C, (unsigned long long)seed, (long long)n, d, 0);
	const cudaError_t e = cudaMemcpy(&hcount, d, sizeof(hcount), cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) return LUMA_B200_ECUDA;
	*mismatches = (int64_t)hcount;
	return LUMA_B200_OK;
}

int luma_b200_sync(luma_b200_t *h)
{
	if (!h) return LUMA_B200_EINVAL;
	CK(cudaSetDevice(h->p.device));
	if (h->have_state)
	{
		const int rc = flush_steps(h);
		if (rc) return rc;
	}
	CK(cudaStreamSynchronize(h->s_main));
	CK(cudaStreamSynchronize(h->s_comm));
	CK(cudaStreamSynchronize(h->s_copy));
	h->copy_pending = false;
	return halo_health(h);
}

}  // extern "C"
