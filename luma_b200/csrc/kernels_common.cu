// kernels_common.cu -- lattice-independent kernels: device-initiated halo exchange, cell-word fix-ups,
// layout conversion of vector fields, eType conversion, the div_const self-test.
#include "kernels_impl.cuh"

namespace luma {

// sites the host wants handled per link (class 4) although they are eFluid
__global__ void k_force_general(uint32_t *cw, const long long *ids, int n, int class_shift)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	const uint32_t w = cw[ids[t]];
	if (((w >> class_shift) & CW_CLASS_MASK) == CLS_FLUID)
		cw[ids[t]] = (w & ~(CW_CLASS_MASK << class_shift)) | (CLS_GENERAL << class_shift);
}
void launch_force_general(uint32_t *cw, const long long *ids, int n, int class_shift, cudaStream_t s)
{
	if (n > 0) k_force_general<<<(n + 127) / 128, 128, 0, s>>>(cw, ids, n, class_shift);
}

// does any site carry one of the link bits in `mask`?  (geometry finalisation: are there walls / bodies bounded in z?)
__global__ void k_any_bits(const uint32_t *__restrict__ cw, long long first, long long n, uint32_t mask, int *flag)
{
	bool hit = false;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
		hit = hit || (cw[first + i] & mask) != 0u;
	if (__any_sync(0xffffffffu, hit) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
void launch_any_bits(const uint32_t *cw, long long first, long long n, uint32_t mask, int *flag, cudaStream_t s)
{
	if (n > 0) k_any_bits<<<148 * 8, 256, 0, s>>>(cw, first, n, mask, flag);
}

__global__ void k_scatter_u32(uint32_t *out, const long long *ids, const uint32_t *vals, int n)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) out[ids[t]] = vals[t];
}
void launch_scatter_u32(uint32_t *out, const long long *ids, const uint32_t *vals, int n, cudaStream_t s)
{
	if (n > 0) k_scatter_u32<<<(n + 127) / 128, 128, 0, s>>>(out, ids, vals, n);
}

// ------------------------------------------------------------------------------------------------
// Halo exchange without a communication library: the populations that leave through a slab face are
// stored straight into the neighbour GPU's ghost plane (NVLink peer memory, mapped through CUDA IPC);
// the last CTA to finish publishes the exchange number in the neighbours' arrival flags
// (fence.sys + release store), and the receiver's k_halo_wait acquires it before anything reads the
// ghost planes.  Replaces MpiManager::mpi_communicate's pack / MPI_Isend / MPI_Recv / unpack
// (src/MpiManager.cpp:631-815) by one store per population element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_halo_push(const HaloPushArgs a)
{
	const int m = blockIdx.y;
	const double *__restrict__ src = a.src[m];
	double *__restrict__ dst = a.dst[m];
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += (long long)gridDim.x * blockDim.x)
		dst[i] = src[i];
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0)
	{
		const unsigned int total = gridDim.x * gridDim.y;
		if (atomicAdd(a.done, 1u) == total - 1)
		{
			*a.done = 0;
			const unsigned long long value = *a.seq + 1;      // this rank's exchange number lives on the device (see k_halo_publish)
			*a.seq = value;
			__threadfence_system();
			for (int side = 0; side < 2; ++side)
				asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(a.peer_flag[side]), "l"(value) : "memory");
		}
	}
}

// the flag part alone, for the fused exchange: the face kernels (k_step_faces, k_bc) of this step have already stored
// the populations into the neighbours' ghost planes and have completed (stream order); the release below is cumulative
// over those stores
// The exchange number is a counter in this rank's own device memory, advanced here, not a value the host passes in: the
// launch arguments are then the same for every step and a batch of steps can be captured into a CUDA graph and replayed.
__global__ void k_halo_publish(unsigned long long *left_flag, unsigned long long *right_flag, unsigned long long *seq)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	const unsigned long long value = *seq + 1;
	*seq = value;
	__threadfence_system();
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(left_flag), "l"(value) : "memory");
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(right_flag), "l"(value) : "memory");
}
void launch_halo_publish(unsigned long long *left_flag, unsigned long long *right_flag, unsigned long long *seq, cudaStream_t s)
{
	k_halo_publish<<<1, 32, 0, s>>>(left_flag, right_flag, seq);
}

void launch_halo_push(const HaloPushArgs &a, cudaStream_t s)
{
	if (a.nmsg <= 0) return;
	unsigned bx = (unsigned)((a.count + 255) / 256);
	if (bx > 64) bx = 64;
	k_halo_push<<<dim3(bx, (unsigned)a.nmsg), 256, 0, s>>>(a);
}

// thread 0 waits for the left neighbour's data, thread 1 for the right neighbour's; gives up after
// 20 s (a dead peer must not hang the GPU) and reports it through *timed_out, a word in mapped host memory that
// the host polls.  Once it is set every later wait returns at once: the queued steps drain in microseconds
// (on stale ghost planes -- the call that notices the flag returns LUMA_B200_ENCCL and the state is void).
__global__ void k_halo_wait(const unsigned long long *flags, const unsigned long long *seq, int *timed_out)
{
	if (threadIdx.x > 1) return;
	if (*reinterpret_cast<volatile int *>(timed_out) != 0) return;
	// the exchange this rank has just published (same stream, earlier kernel): its neighbours' data of the same number must be in
	const unsigned long long value = *reinterpret_cast<const volatile unsigned long long *>(seq);
	unsigned long long t0, t1, seen;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for (;;)
	{
		asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + threadIdx.x) : "memory");
		if (seen >= value) break;
		__nanosleep(64);
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (t1 - t0 > 20000000000ull) { *reinterpret_cast<volatile int *>(timed_out) = 1; __threadfence_system(); break; }
	}
}

void launch_halo_wait(const unsigned long long *flags, const unsigned long long *seq, int *timed_out, cudaStream_t s)
{
	k_halo_wait<<<1, 32, 0, s>>>(flags, seq, timed_out);
}

void launch_u_aos_to_soa(const double *aos, double *soa, long long stride, int ncomp, long long first, long long n, cudaStream_t s)
{
	if (n <= 0) return;
	if (ncomp == 6) k_aos_to_soa<6><<<(unsigned)((n + 63) / 64), 192, 0, s>>>(aos, soa, stride, first, n);
	else if (ncomp == 3) k_aos_to_soa<3><<<(unsigned)((n + 63) / 64), 192, 0, s>>>(aos, soa, stride, first, n);
	else k_aos_to_soa<2><<<(unsigned)((n + 63) / 64), 128, 0, s>>>(aos, soa, stride, first, n);
}
void launch_u_soa_to_aos(const double *soa, double *aos, long long stride, int ncomp, long long first, long long n, cudaStream_t s)
{
	if (n <= 0) return;
	if (ncomp == 6) k_soa_to_aos<6><<<(unsigned)((n + 63) / 64), 192, 0, s>>>(soa, aos, stride, first, n);
	else if (ncomp == 3) k_soa_to_aos<3><<<(unsigned)((n + 63) / 64), 192, 0, s>>>(soa, aos, stride, first, n);
	else k_soa_to_aos<2><<<(unsigned)((n + 63) / 64), 128, 0, s>>>(soa, aos, stride, first, n);
}

__global__ void k_types_from_i32(const int32_t *__restrict__ in, uint8_t *__restrict__ out, long long n)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = (uint8_t)in[i];
}
__global__ void k_types_to_i32(const uint8_t *__restrict__ in, int32_t *__restrict__ out, long long n)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = (int32_t)in[i];
}
void launch_types_from_i32(const int32_t *in, uint8_t *out, long long n, cudaStream_t s)
{
	if (n > 0) k_types_from_i32<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
}
void launch_types_to_i32(const uint8_t *in, int32_t *out, long long n, cudaStream_t s)
{
	if (n > 0) k_types_to_i32<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n);
}

// ------------------------------------------------------------------------------------------------
// self-test: div_const against IEEE division on pseudo-random operands (all binades the step sees)
// ------------------------------------------------------------------------------------------------
__global__ void k_selftest_div(const LbmConst C, unsigned long long seed, long long n, unsigned long long *mismatches)
{
	unsigned long long bad = 0;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
	{
		unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);   // splitmix64
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		z ^= z >> 31;
		const unsigned long long mant = z & 0xFFFFFFFFFFFFFull;
		const unsigned long long expo = 1023ull - 90ull + ((z >> 52) % 100ull);                // 2^-90 .. 2^9
		const unsigned long long sign = (z >> 63) << 63;
		const double a = __longlong_as_double((long long)(sign | (expo << 52) | mant));
		if (div_const(a, C.cs2, C.inv_cs2) != a / C.cs2) ++bad;
		if (div_const(a, C.den, C.inv_den) != a / C.den) ++bad;
	}
	if (bad) atomicAdd(mismatches, bad);
}
void launch_selftest_div(const LbmConst &C, unsigned long long seed, long long n, unsigned long long *mismatches, cudaStream_t s)
{
	k_selftest_div<<<148 * 8, 256, 0, s>>>(C, seed, n, mismatches);
}

}  // namespace luma
