// mps_comm.cu — multi-GPU: one process per GPU, coupled over NVLink 5 / NVSwitch peer memory (NCCL for set-up and as a fallback).
//
// Decomposition (DESIGN.md "multi-GPU"): the cell-sorted slot order is x-major, so a contiguous range of slots IS an
// x-slab of the domain.  Every step the slabs are re-cut between cell columns so that every rank gets an equal share of a
// per-type work model (k_slab_bounds, mps_grid.cu) whatever the shape of the fluid (the dam break starts in the left
// quarter of the tank; equal-width slabs would idle most GPUs — SURVEY H9).  Every rank keeps the whole particle state (a few
// hundred bytes per particle) but COMPUTES only its slab: neighbour lists, gather stages, PPE rows and the CG rows of its
// slots; what grows with the problem — neighbour list and matrix — is therefore partitioned.  Exchanges, all on the
// solver's stream (NCCL itself only for set-up, the per-step sum of the first replicated coarse operator and the halo extents):
//   * peer memory (mps_comm_mode 1): every rank exports one arena (mailboxes, barrier flags, the solve's {z, p} buffers, the
//     level vectors of the preconditioner) and its state arrays through CUDA IPC.  After the stages that move particles or set
//     pressures every rank PULLS the halo (one cell column of each adjacent rank: a contiguous slot range) straight from the
//     owners between two flag barriers (k_peer_barrier, k_peer_gather); one full gather of x, u, p, n ends the step.
//     Re-sorting the replicated state every step replaces particle migration.  The solve is ONE persistent kernel per rank
//     (mps_cg.cu): halo of {z, p} pulled once per iteration, dot products through flag-in-data mailboxes.
//   * NCCL (mode 2: the ranks cannot map each other's memory, or MPS_COMM_NCCL_ONLY=1): in-place ncclAllGather of the
//     fields, a stepwise plain CG with ncclSend/ncclRecv of the rim and ncclAllReduce of the dot products.
// The reference has nothing to compare with here (single process, OpenMP); parity is 1 GPU vs N GPUs on the same input.
//
// NCCL is loaded with dlopen so that the library neither links against a particular libnccl nor fights the copy that
// PyTorch bundles when both live in one process.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nccl.h>

#include "mps_solver.h"

namespace mps {
namespace {

struct Nccl
{
	void* lib = nullptr;
	decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
	decltype(&ncclCommInitRank) CommInitRank = nullptr;
	decltype(&ncclCommDestroy) CommDestroy = nullptr;
	decltype(&ncclAllReduce) AllReduce = nullptr;
	decltype(&ncclAllGather) AllGather = nullptr;
	decltype(&ncclBroadcast) Broadcast = nullptr;
	decltype(&ncclSend) Send = nullptr;
	decltype(&ncclRecv) Recv = nullptr;
	decltype(&ncclGroupStart) GroupStart = nullptr;
	decltype(&ncclGroupEnd) GroupEnd = nullptr;
	decltype(&ncclGetErrorString) GetErrorString = nullptr;
	std::string error;

	bool load()
	{
		if (lib) return true;
		for (const char* name : { "libnccl.so.2", "libnccl.so" })
		{
			lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (lib) break;
		}
		if (!lib) { error = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define MPS_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(lib, #sym)); if (!field) { error = "libnccl lacks " #sym; return false; }
		MPS_SYM(GetUniqueId, ncclGetUniqueId) MPS_SYM(CommInitRank, ncclCommInitRank) MPS_SYM(CommDestroy, ncclCommDestroy)
		MPS_SYM(AllReduce, ncclAllReduce) MPS_SYM(AllGather, ncclAllGather) MPS_SYM(Broadcast, ncclBroadcast) MPS_SYM(Send, ncclSend) MPS_SYM(Recv, ncclRecv)
		MPS_SYM(GroupStart, ncclGroupStart) MPS_SYM(GroupEnd, ncclGroupEnd) MPS_SYM(GetErrorString, ncclGetErrorString)
#undef MPS_SYM
		return true;
	}
};

Nccl g_nccl;

struct PeerBarrierArgs
{
	int rank, nranks;
	unsigned long long seq;
	unsigned long long* flags[kMaxPeerRanks]; // every rank's arrival counters (peer-mapped), [r][k] = arrivals of rank k seen by rank r
};
struct PeerGatherArgs
{
	int rank, nranks, nfields;
	unsigned long long* dst[4];
	const unsigned long long* src[4][kMaxPeerRanks];
	unsigned long long words_per_row[4];
	unsigned long long lo[kMaxPeerRanks], hi[kMaxPeerRanks]; // rows [lo, hi) to pull from each rank (its whole slab, or the part of it in this rank's halo)
};

#define MPS_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)
#define MPS_NCCL(s, expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) { (s)->comm_error = std::string(#expr ": ") + g_nccl.GetErrorString(r_); return cudaErrorUnknown; } } while (0)

inline ncclComm_t comm_of(mps_solver* s) { return static_cast<ncclComm_t>(s->comm.nccl); }

// extent of the slots this rank's windows reach: min range start / max range end over its chunks with entries
__global__ void k_halo_extent(const ChunkDesc* __restrict__ desc, const DevScalars* __restrict__ sc, unsigned long long* ext)
{
	const unsigned long long nchunks = sc->n_chunks;
	unsigned long long lo = ~0ull, hi = 0;
	for (unsigned long long c = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; c < nchunks; c += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
	{
		const ChunkDesc& d = desc[c];
		if (d.nnz == 0) continue;
		for (uint32_t q = 0; q < d.nranges; q++)
		{
			const unsigned long long b = d.range_start[q], e = b + d.range_len[q];
			if (b < lo) lo = b;
			if (e > hi) hi = e;
		}
	}
	if (lo != ~0ull) { atomicMin(&ext[0], lo); atomicMax(&ext[1], hi); }
}

struct Slab { uint64_t b, e; };
inline Slab slab_of(const mps_solver* s, int rank) { return Slab{ s->slab_begin(rank), s->slab_begin(rank + 1) }; }

// Halo extents of every rank (once per solve: the chunks were rebuilt by this step's assembly): ext[2k], ext[2k+1] = first /
// one-past-last slot the windows of rank k reach.  One small all-gather + one host read-back.
cudaError_t halo_extents(mps_solver* s, std::vector<unsigned long long>& ext)
{
	CgBuffers& c = s->cg;
	cudaStream_t st = s->stream;
	const int R = s->comm.nranks;
	MPS_TRY(s->comm.ext.ensure(2ull * R + 2, st));
	unsigned long long* ext_local = s->comm.ext.p + 2ull * R;
	const unsigned long long init[2] = { s->own0(), s->own1() };
	MPS_TRY(cudaMemcpyAsync(ext_local, init, sizeof(init), cudaMemcpyHostToDevice, st));
	k_halo_extent<<<64, 256, 0, st>>>(c.desc.p, s->d_sc, ext_local);
	s->stats.kernel_launches += 1;
	MPS_NCCL(s, g_nccl.AllGather(ext_local, s->comm.ext.p, 2, ncclUint64, comm_of(s), st));
	s->stats.comm_calls += 1;
	ext.assign(2ull * R, 0);
	MPS_TRY(cudaMemcpyAsync(ext.data(), s->comm.ext.p, 2ull * R * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	MPS_TRY(cudaStreamSynchronize(st));
	for (int k = 0; k < R; k++)
	{
		// a window must not reach past the adjacent slab (slabs thinner than one cell column are not supported)
		const uint64_t lo_ok = (k > 0) ? s->slab_begin(k - 1) : 0;
		const uint64_t hi_ok = (k + 1 < R) ? s->slab_begin(k + 2) : s->n;
		if (ext[2 * k] < lo_ok || ext[2 * k + 1] > hi_ok + 1 /* window ends are rounded up to even */) { s->comm_error = "slab thinner than the neighbour stencil: use fewer GPUs for this problem"; return cudaErrorUnknown; }
	}
	return cudaSuccess;
}

// rim exchange of one {r, p} buffer with the two adjacent ranks (host copies of all extents in `ext`)
cudaError_t halo_exchange(mps_solver* s, double* z, const std::vector<unsigned long long>& ext)
{
	const int k = s->comm.rank, R = s->comm.nranks;
	const Slab me = slab_of(s, k);
	auto clampu = [](uint64_t v, uint64_t lo, uint64_t hi) { return v < lo ? lo : (v > hi ? hi : v); };
	MPS_NCCL(s, g_nccl.GroupStart());
	if (k > 0)
	{
		const Slab left = slab_of(s, k - 1);
		// the left rank reads my rows [me.b, ext_hi(left)); I read its rows [ext_lo(me), me.b)
		const uint64_t send_e = clampu(ext[2 * (k - 1) + 1], me.b, me.e);
		if (send_e > me.b) MPS_NCCL(s, g_nccl.Send(z + 2 * me.b, 2 * (send_e - me.b), ncclDouble, k - 1, comm_of(s), s->stream));
		const uint64_t recv_b = clampu(ext[2 * k], left.b, me.b);
		if (recv_b < me.b) MPS_NCCL(s, g_nccl.Recv(z + 2 * recv_b, 2 * (me.b - recv_b), ncclDouble, k - 1, comm_of(s), s->stream));
	}
	if (k + 1 < R)
	{
		const Slab right = slab_of(s, k + 1);
		const uint64_t send_b = clampu(ext[2 * (k + 1)], me.b, me.e);
		if (send_b < me.e) MPS_NCCL(s, g_nccl.Send(z + 2 * send_b, 2 * (me.e - send_b), ncclDouble, k + 1, comm_of(s), s->stream));
		const uint64_t recv_e = clampu(ext[2 * k + 1], me.e, right.e);
		if (recv_e > me.e) MPS_NCCL(s, g_nccl.Recv(z + 2 * me.e, 2 * (recv_e - me.e), ncclDouble, k + 1, comm_of(s), s->stream));
	}
	MPS_NCCL(s, g_nccl.GroupEnd());
	return cudaSuccess;
}

} // namespace

// ---- peer-memory coupling of the persistent CG kernel ----------------------------------------------------------------------
namespace {

inline size_t arena_mg_offset(uint64_t rows) { return kPeerHeaderBytes + 2ull * rows * sizeof(double2); } // [header | z0 | z1 | mg]
inline size_t arena_size(uint64_t rows, uint64_t mg_doubles) { return arena_mg_offset(rows) + mg_doubles * sizeof(double); }

bool allocation_offset(void* p, unsigned long long* offset);

struct PeerExport
{
	cudaIpcMemHandle_t handle; // of the allocation that contains the buffer
	unsigned long long offset; // buffer - allocation base (cudaMalloc may sub-allocate small requests)
	unsigned long long bytes;
	unsigned long long ok;     // 0: this rank could not export
};
constexpr int kExports = 9;    // the arena + pos[2], vel[2], prs[2], nden[2]

inline void state_pointers(mps_solver* s, const void* out[8])
{
	for (int b = 0; b < 2; b++) { out[0 + b] = s->pos[b].p; out[2 + b] = s->vel[b].p; out[4 + b] = s->prs[b].p; out[6 + b] = s->nden[b].p; }
}

bool export_buffer(void* p, PeerExport& e)
{
	e.ok = 0;
	if (!p) return false;
	if (!allocation_offset(p, &e.offset)) return false;
	if (cudaIpcGetMemHandle(&e.handle, static_cast<unsigned char*>(p) - e.offset) != cudaSuccess) { cudaGetLastError(); return false; }
	e.ok = 1;
	return true;
}

// offset of `p` inside its allocation (the IPC handle always opens the allocation's base)
bool allocation_offset(void* p, unsigned long long* offset)
{
	typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
	static GetRange fn = nullptr;
	if (!fn)
	{
		void* sym = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return false;
		fn = reinterpret_cast<GetRange>(sym);
	}
	unsigned long long base = 0; size_t size = 0;
	if (fn(&base, &size, reinterpret_cast<unsigned long long>(p)) != 0) return false;
	*offset = reinterpret_cast<unsigned long long>(p) - base;
	return true;
}

// every rank's stream reaches this point before anybody continues (used before exported memory is freed)
cudaError_t comm_barrier(mps_solver* s)
{
	MPS_TRY(s->comm.ext.ensure(2ull * s->comm.nranks + 2, s->stream));
	MPS_NCCL(s, g_nccl.AllReduce(s->comm.ext.p, s->comm.ext.p, 1, ncclUint64, ncclSum, comm_of(s), s->stream));
	return cudaStreamSynchronize(s->stream);
}

void close_peers(mps_solver* s)
{
	Comm& c = s->comm;
	for (int r = 0; r < kMaxPeerRanks; r++)
	{
		if (c.peer_base[r]) cudaIpcCloseMemHandle(c.peer_base[r]);
		c.peer_base[r] = nullptr; c.peer_arena[r] = nullptr;
		for (int f = 0; f < 8; f++)
		{
			if (c.peer_state_base[r][f]) cudaIpcCloseMemHandle(c.peer_state_base[r][f]);
			c.peer_state_base[r][f] = nullptr; c.peer_state[r][f] = nullptr;
		}
	}
}

} // namespace

void comm_release_peers(mps_solver* s)
{
	Comm& c = s->comm;
	if (!c.arena) return;
	close_peers(s);
	if (c.on && c.peer_mode == 1) comm_barrier(s); // nobody frees memory a peer still maps
	cudaFree(c.arena);
	c.arena = nullptr; c.arena_bytes = 0; c.arena_rows = 0; c.arena_mg = 0;
	if (s->cg.z_borrowed) { s->cg.z0.p = nullptr; s->cg.z0.cap = 0; s->cg.z1.p = nullptr; s->cg.z1.cap = 0; s->cg.z_borrowed = false; }
}

// What the other ranks read of this one, shared through CUDA IPC: the arena for `rows` rows (mailbox, barrier flags, {r, p}
// buffers of the CG kernel) and the replicated state arrays.  (Re)done when the arena is too small or a state array was
// re-allocated — every rank holds the same particles, so all of them take this branch in the same step.  IPC handles travel
// through an NCCL all-gather of a small device buffer; peers map them with cudaIpcOpenMemHandle (which enables peer access).
// If any rank cannot export or map (no P2P, IPC disabled in the container) all ranks agree on the NCCL paths.
cudaError_t comm_ensure_arena(mps_solver* s, uint64_t rows, uint64_t mg_doubles)
{
	Comm& c = s->comm;
	CgBuffers& cg = s->cg;
	cudaStream_t st = s->stream;
	const void* now[8];
	state_pointers(s, now);
	bool same = c.arena && rows <= c.arena_rows && mg_doubles <= c.arena_mg;
	for (int f = 0; f < 8 && same; f++) same = (now[f] == c.exported[f]);
	if (same) return cudaSuccess;
	const int R = c.nranks;
	const bool regrow = !c.arena || rows > c.arena_rows || mg_doubles > c.arena_mg;
	const uint64_t old_rows = c.arena_rows, old_mg = c.arena_mg; // a regrown arena never shrinks either section
	MPS_TRY(cudaStreamSynchronize(st));
	if (c.arena && regrow) comm_release_peers(s);
	else if (c.arena) { close_peers(s); if (c.peer_mode == 1) MPS_TRY(comm_barrier(s)); }
	if (regrow)
	{
		if (!cg.z_borrowed) { cg.z0.release(); cg.z1.release(); }
		const uint64_t want_rows = std::max<uint64_t>(rows + rows / 4 + 16, old_rows);
		const uint64_t want_mg = std::max<uint64_t>(mg_doubles ? mg_doubles + mg_doubles / 4 + 64 : 0, old_mg);
		size_t bytes = arena_size(want_rows, want_mg);
		bytes = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
		void* mem = nullptr;
		MPS_TRY(cudaMalloc(&mem, bytes));
		MPS_TRY(cudaMemsetAsync(mem, 0, bytes, st));
		c.arena = static_cast<unsigned char*>(mem); c.arena_bytes = bytes; c.arena_rows = want_rows; c.arena_mg = want_mg;
		cg.z0.p = reinterpret_cast<double*>(c.arena + kPeerHeaderBytes); cg.z0.cap = 2 * want_rows;
		cg.z1.p = cg.z0.p + 2 * want_rows; cg.z1.cap = 2 * want_rows;
		cg.z_borrowed = true;
		c.bar_seq = 0; // the new arena's flags start from zero on every rank
	}
	for (int f = 0; f < 8; f++) c.exported[f] = now[f];

	PeerExport mine[kExports] = {};
	const bool want_peer = (R <= kMaxPeerRanks) && !std::getenv("MPS_COMM_NCCL_ONLY");
	if (want_peer)
	{
		export_buffer(c.arena, mine[0]);
		for (int f = 0; f < 8; f++) export_buffer(const_cast<void*>(now[f]), mine[1 + f]);
	}
	cudaGetLastError(); // a failed export is not an error of the solver: it selects the NCCL path
	static_assert(sizeof(PeerExport) % 8 == 0, "exchanged as 64-bit words");
	DevBuf<unsigned long long> xchg;
	const size_t words = sizeof(mine) / 8;
	MPS_TRY(xchg.ensure(words * (R + 1), st));
	MPS_TRY(cudaMemcpyAsync(xchg.p + words * R, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
	MPS_NCCL(s, g_nccl.AllGather(xchg.p + words * R, xchg.p, words, ncclUint64, comm_of(s), st));
	std::vector<PeerExport> all(static_cast<size_t>(R) * kExports);
	MPS_TRY(cudaMemcpyAsync(all.data(), xchg.p, sizeof(mine) * R, cudaMemcpyDeviceToHost, st));
	MPS_TRY(cudaStreamSynchronize(st));
	xchg.release();
	s->stats.comm_calls += 1;

	unsigned long long ok = 1;
	for (size_t k = 0; k < all.size(); k++) ok &= all[k].ok;
	for (int r = 0; r < R && ok; r++)
	{
		const PeerExport* e = &all[static_cast<size_t>(r) * kExports];
		if (r == c.rank)
		{
			c.peer_arena[r] = c.arena;
			for (int f = 0; f < 8; f++) c.peer_state[r][f] = static_cast<unsigned char*>(const_cast<void*>(now[f]));
			continue;
		}
		// one allocation may hold several of the buffers (cudaMalloc sub-allocates small requests): open every handle once
		void* opened[kExports] = {};
		for (int k = 0; k < kExports && ok; k++)
		{
			void* base = nullptr;
			for (int j = 0; j < k; j++) if (std::memcmp(&e[j].handle, &e[k].handle, sizeof(cudaIpcMemHandle_t)) == 0) { base = opened[j]; break; }
			if (!base)
			{
				if (cudaIpcOpenMemHandle(&base, e[k].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
				if (k == 0) c.peer_base[r] = base; else c.peer_state_base[r][k - 1] = base;
			}
			opened[k] = base;
			unsigned char* ptr = static_cast<unsigned char*>(base) + e[k].offset;
			if (k == 0) c.peer_arena[r] = ptr; else c.peer_state[r][k - 1] = ptr;
		}
	}
	// agree: one rank that cannot map sends everybody to the NCCL path
	MPS_TRY(c.ext.ensure(2ull * R + 2, st));
	unsigned long long* flag = c.ext.p;
	MPS_TRY(cudaMemcpyAsync(flag, &ok, sizeof(ok), cudaMemcpyHostToDevice, st));
	MPS_NCCL(s, g_nccl.AllReduce(flag, flag, 1, ncclUint64, ncclMin, comm_of(s), st));
	MPS_TRY(cudaMemcpyAsync(&ok, flag, sizeof(ok), cudaMemcpyDeviceToHost, st));
	MPS_TRY(cudaStreamSynchronize(st));
	s->stats.comm_calls += 1;
	if (!ok) close_peers(s);
	c.peer_mode = ok ? 1 : 2;
	return cudaSuccess;
}

// ---- the per-stage all-gather of replicated fields over peer memory ------------------------------------------------------
namespace {

// Stream-level barrier across the ranks (one warp): my arrival number into every rank's flag array, then wait for everybody's.
// Launched between kernels: what the previous kernels of this stream wrote is complete and sits in this GPU's L2.
__global__ void k_peer_barrier(PeerBarrierArgs a)
{
	const unsigned lane = threadIdx.x;
	if (lane < static_cast<unsigned>(a.nranks))
	{
		asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flags[lane] + a.rank), "l"(a.seq) : "memory");
		unsigned long long v = 0, t0 = 0, t1 = 0;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		for (unsigned spin = 0;; spin++)
		{
			asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags[a.rank] + lane) : "memory");
			if (v >= a.seq) break;
			if ((spin & 1023u) == 1023u)
			{
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
				if (t1 - t0 > 30ull * 1000ull * 1000ull * 1000ull) __trap(); // a dead rank ends as a failed launch, not as a hung GPU
			}
		}
		asm volatile("fence.acq_rel.sys;" ::: "memory");
	}
}

// every rank pulls the other ranks' slabs of up to four fields straight from their memory: 16-byte words (coalesced, four in
// flight per thread: NVLink wants many outstanding requests), 8-byte ends where a range starts or ends on an odd word
__global__ void __launch_bounds__(256) k_peer_gather(PeerGatherArgs a)
{
	const unsigned long long tid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
	const unsigned long long nthreads = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
	for (int f = 0; f < a.nfields; f++)
	{
		const unsigned long long w = a.words_per_row[f];
		unsigned long long* __restrict__ dst = a.dst[f];
		for (int r = 0; r < a.nranks; r++)
		{
			if (r == a.rank) continue;
			unsigned long long b = a.lo[r] * w, e = a.hi[r] * w;
			if (e <= b) continue;
			const unsigned long long* __restrict__ src = a.src[f][r];
			// odd ends (the arrays are 16-byte aligned, so word parity = address parity)
			if (b & 1ull) { if (tid == 0) dst[b] = src[b]; b++; }
			if (e & 1ull) { if (tid == 0) dst[e - 1] = src[e - 1]; e--; }
			const ulonglong2* __restrict__ s2 = reinterpret_cast<const ulonglong2*>(src + b);
			ulonglong2* __restrict__ d2 = reinterpret_cast<ulonglong2*>(dst + b);
			const unsigned long long m = (e - b) >> 1;
			for (unsigned long long i = tid; i < m; i += 4 * nthreads)
			{
				ulonglong2 v[4];
#pragma unroll
				for (int u = 0; u < 4; u++) { const unsigned long long j = i + u * nthreads; if (j < m) v[u] = s2[j]; }
#pragma unroll
				for (int u = 0; u < 4; u++) { const unsigned long long j = i + u * nthreads; if (j < m) d2[j] = v[u]; }
			}
		}
	}
}

cudaError_t peer_barrier(mps_solver* s)
{
	Comm& c = s->comm;
	PeerBarrierArgs a{};
	a.rank = c.rank; a.nranks = c.nranks; a.seq = ++c.bar_seq;
	for (int r = 0; r < c.nranks; r++) a.flags[r] = reinterpret_cast<unsigned long long*>(c.peer_arena[r] + kPeerBarrierOff);
	k_peer_barrier<<<1, 32, 0, s->stream>>>(a);
	s->stats.kernel_launches += 1;
	return cudaGetLastError();
}

} // namespace

// Fields that neighbours read, after the stage that wrote this rank's slab of them.  Peer memory: barrier (every rank's slab
// is written), one kernel in which every rank copies the other slabs straight from their owners over NVLink, barrier (nobody
// overwrites a slab a peer is still reading).  NCCL all-gathers where the ranks cannot map each other's memory.
// Inside a time step (comm.halo_step) only the HALO is pulled: the stages of a step read other particles through the neighbour
// list, whose candidates lie in the 3^D cell stencil, i.e. at most one cell column beyond this rank's slab on either side — one
// contiguous range of the cell-sorted slots per adjacent rank (halo_lo / halo_hi, k_slab_bounds).  The step ends with one full
// gather of x, u, p, n (mps_capi.cu step_body): the next sort, DetermineDt and Particles() see the whole replicated state.
cudaError_t comm_allgather_state(mps_solver* s, bool pos, bool vel, bool prs, bool nden)
{
	if (!s->comm.on || s->n == 0) return cudaSuccess;
	Comm& c = s->comm;
	const int vs = s->vec_stride();
	const int k = c.rank;
	MPS_TRY(comm_ensure_arena(s, s->n + 64)); // first use (or re-allocated state): exchange the IPC handles
	if (c.peer_mode == 1)
	{
		PeerGatherArgs a{};
		a.rank = k; a.nranks = c.nranks;
		auto add = [&](int field, uint64_t words_per_row)
		{
			const int f = 2 * field + s->cur, q = a.nfields++;
			a.dst[q] = reinterpret_cast<unsigned long long*>(c.peer_state[k][f]);
			for (int r = 0; r < c.nranks; r++) a.src[q][r] = reinterpret_cast<const unsigned long long*>(c.peer_state[r][f]);
			a.words_per_row[q] = words_per_row;
		};
		const bool halo = c.halo_step && s->slabs_set() && s->halo_lo.size() == static_cast<size_t>(c.nranks);
		for (int r = 0; r < c.nranks; r++)
		{
			uint64_t lo = s->slab_begin(r), hi = s->slab_begin(r + 1);
			if (halo)
			{
				// my halo: [halo_lo[k], own0) in the slab of rank k - 1 and [own1, halo_hi[k]) in the slab of rank k + 1
				if (r == k - 1) lo = std::max<uint64_t>(lo, s->halo_lo[k]);
				else if (r == k + 1) hi = std::min<uint64_t>(hi, s->halo_hi[k]);
				else hi = lo;
			}
			a.lo[r] = lo; a.hi[r] = (hi > lo) ? hi : lo;
		}
		if (pos) add(0, vs);
		if (vel) add(1, vs);
		if (prs) add(2, 1);
		if (nden) add(3, 1);
		if (a.nfields == 0) return cudaSuccess;
		MPS_TRY(peer_barrier(s));
		k_peer_gather<<<4 * s->sm_count, 256, 0, s->stream>>>(a);
		s->stats.kernel_launches += 1;
		MPS_TRY(peer_barrier(s));
		c.peer_gathers += 1;
		return cudaGetLastError();
	}
	// slabs are uneven (cell-column aligned): every rank broadcasts its slab, one group per field
	auto gather = [&](double* p, uint64_t w) -> cudaError_t
	{
		MPS_NCCL(s, g_nccl.GroupStart());
		for (int r = 0; r < c.nranks; r++)
		{
			const uint64_t b = s->slab_begin(r), e = s->slab_begin(r + 1);
			if (e > b) MPS_NCCL(s, g_nccl.Broadcast(p + b * w, p + b * w, (e - b) * w, ncclDouble, r, comm_of(s), s->stream));
		}
		MPS_NCCL(s, g_nccl.GroupEnd());
		return cudaSuccess;
	};
	if (pos) MPS_TRY(gather(s->pos[s->cur].p, vs));
	if (vel) MPS_TRY(gather(s->vel[s->cur].p, vs));
	if (prs) MPS_TRY(gather(s->prs[s->cur].p, 1));
	if (nden) MPS_TRY(gather(s->nden[s->cur].p, 1));
	s->stats.comm_calls += (pos ? 1 : 0) + (vel ? 1 : 0) + (prs ? 1 : 0) + (nden ? 1 : 0);
	return cudaSuccess;
}

double* comm_mg_section(mps_solver* s, int rank)
{
	Comm& c = s->comm;
	if (!c.peer_arena[rank]) return nullptr;
	return reinterpret_cast<double*>(c.peer_arena[rank] + arena_mg_offset(c.arena_rows));
}

cudaError_t comm_allreduce_sum(mps_solver* s, double* p, uint64_t count)
{
	if (!s->comm.on || count == 0) return cudaSuccess;
	MPS_NCCL(s, g_nccl.AllReduce(p, p, count, ncclDouble, ncclSum, comm_of(s), s->stream));
	s->stats.comm_calls += 1;
	return cudaSuccess;
}

// comm.link for the next persistent solve: neighbours' {r, p} buffers, our rows that they read, everybody's mailbox
cudaError_t comm_prepare_link(mps_solver* s)
{
	Comm& c = s->comm;
	const int k = c.rank, R = c.nranks;
	std::vector<unsigned long long> ext;
	MPS_TRY(halo_extents(s, ext));
	PeerLink& L = c.link;
	L = PeerLink{};
	L.rank = k; L.nranks = R;
	c.solves += 1;
	L.tag = c.solves << 32;
	const size_t z1_off = kPeerHeaderBytes + 2ull * c.arena_rows * sizeof(double);
	for (int side = 0; side < 2; side++)
	{
		const int nb = (side == 0) ? k - 1 : k + 1;
		if (nb < 0 || nb >= R) continue;
		L.nb_z0[side] = reinterpret_cast<double2*>(c.peer_arena[nb] + kPeerHeaderBytes);
		L.nb_z1[side] = reinterpret_cast<double2*>(c.peer_arena[nb] + z1_off);
	}
	for (int r = 0; r < R; r++) L.mail[r] = reinterpret_cast<PeerMail*>(c.peer_arena[r]);
	return cudaSuccess;
}

// Computer::SolvePressurePoissonEquation (Computer.hpp:1359-1429) across ranks: the same CG, the same stopping rule; one
// launch per phase so that the dot products can be all-reduced and the rim of {r, p} exchanged between them.
cudaError_t comm_cg_solve(mps_solver* s)
{
	CgBuffers& c = s->cg;
	cudaStream_t st = s->stream;
	MPS_TRY(c.step.ensure(1, st));
	if (!c.h_step) MPS_TRY(cudaMallocHost(&c.h_step, sizeof(CgStepScalars)));
	MPS_TRY(cudaMemsetAsync(c.step.p, 0, sizeof(CgStepScalars), st));

	std::vector<unsigned long long> ext;
	MPS_TRY(halo_extents(s, ext));

	double* z0 = c.z0.p;
	double* z1 = c.z1.p;
	CgStepScalars* d = c.step.p;
	// r0 = b - A x
	MPS_TRY(launch_cg_step(s, 0, d));
	MPS_TRY(launch_cg_reduce(s, &d->rr));
	MPS_NCCL(s, g_nccl.AllReduce(&d->rr, &d->rr, 1, ncclDouble, ncclSum, comm_of(s), st));
	MPS_TRY(launch_cg_scalars(s, 0, d));
	MPS_TRY(halo_exchange(s, z0, ext));
	int zcur_is_1 = 1;
	const uint64_t max_iter = c.n;
	uint64_t iter = 0;
	for (;;)
	{
		MPS_TRY(cudaMemcpyAsync(c.h_step, d, sizeof(CgStepScalars), cudaMemcpyDeviceToHost, st));
		MPS_TRY(cudaStreamSynchronize(st));
		if (c.h_step->converged || iter >= max_iter) break;
		MPS_TRY(launch_cg_step(s, 1, d));
		MPS_TRY(launch_cg_reduce(s, &d->pAp));
		MPS_NCCL(s, g_nccl.AllReduce(&d->pAp, &d->pAp, 1, ncclDouble, ncclSum, comm_of(s), st));
		MPS_TRY(launch_cg_step(s, 2, d));
		MPS_TRY(launch_cg_reduce(s, &d->rr_new));
		MPS_NCCL(s, g_nccl.AllReduce(&d->rr_new, &d->rr_new, 1, ncclDouble, ncclSum, comm_of(s), st));
		MPS_TRY(launch_cg_scalars(s, 1, d));
		MPS_TRY(halo_exchange(s, zcur_is_1 ? z1 : z0, ext)); // the buffer this iteration wrote is the next "previous"
		zcur_is_1 ^= 1;
		iter++;
		s->stats.comm_calls += 4;
	}
	MPS_TRY(launch_cg_scalars(s, 2, d));
	return cudaGetLastError();
}

} // namespace mps

using namespace mps;

extern "C" {

int mps_comm_unique_id(void* out128)
{
	if (!out128) return MPS_BAD_ARG;
	if (!g_nccl.load()) return MPS_NCCL_ERROR;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MPS_NCCL_ERROR;
	std::memcpy(out128, &id, sizeof(id));
	return MPS_OK;
}

int mps_comm_init(mps_handle s, int rank, int nranks, const void* id128)
{
	if (!s || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return MPS_BAD_ARG;
	if (s->n != 0 && s->searched) { s->last_error = "attach the communicator before the first step"; return MPS_BAD_ARG; }
	if (nranks == 1) { s->comm.on = false; return MPS_OK; }
	if (!g_nccl.load()) { s->last_error = g_nccl.error; return MPS_NCCL_ERROR; }
	cudaSetDevice(s->device);
	ncclUniqueId id;
	std::memcpy(&id, id128, sizeof(id));
	ncclComm_t comm = nullptr;
	const ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
	if (r != ncclSuccess) { s->last_error = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return MPS_NCCL_ERROR; }
	s->comm.nccl = comm; s->comm.rank = rank; s->comm.nranks = nranks; s->comm.on = true;
	return MPS_OK;
}

int mps_comm_info(mps_handle s, int* rank, int* nranks, uint64_t* own_first, uint64_t* own_last)
{
	if (!s) return MPS_BAD_ARG;
	if (rank) *rank = s->comm.rank;
	if (nranks) *nranks = s->comm.on ? s->comm.nranks : 1;
	if (own_first) *own_first = s->own0();
	if (own_last) *own_last = s->own1();
	return MPS_OK;
}

int mps_comm_mode(mps_handle s, int* mode)
{
	if (!s || !mode) return MPS_BAD_ARG;
	*mode = s->comm.on ? s->comm.peer_mode : 0;
	return MPS_OK;
}

// pure arithmetic of the slab decomposition (no GPU needed): slots [first, last) of `rank` among `nranks` for n particles
int mps_partition_range(uint64_t n, int nranks, int rank, uint64_t* first, uint64_t* last)
{
	if (nranks < 1 || rank < 0 || rank >= nranks || !first || !last) return MPS_BAD_ARG;
	const uint64_t m = (n + static_cast<uint64_t>(nranks) - 1) / static_cast<uint64_t>(nranks);
	uint64_t b = static_cast<uint64_t>(rank) * m, e = b + m;
	if (b > n) b = n;
	if (e > n) e = n;
	*first = b; *last = e;
	return MPS_OK;
}

void mps_comm_release(mps_handle s)
{
	if (s && s->comm.nccl && g_nccl.CommDestroy) { g_nccl.CommDestroy(static_cast<ncclComm_t>(s->comm.nccl)); s->comm.nccl = nullptr; s->comm.on = false; }
}

} // extern "C"
