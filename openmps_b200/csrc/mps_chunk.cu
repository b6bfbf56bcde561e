// mps_chunk.cu — cuts the rows of the pressure Poisson equation into "chunks" whose matrix entries AND gathered-vector
// window fit the shared-memory stages of the CG kernel (mps_device.cuh "chunk blob", DESIGN.md "CG kernel").
//
// This replaces what the reference does serially on the host after assembling the rows — inserting them into a uBLAS
// compressed_matrix and copying that into a ViennaCL CSR (Computer.hpp:1337-1354, viennacl/compressed_matrix.hpp:164-191):
// here the "matrix format conversion" is a layout decision made on the device before the rows are written.
//
// One thread partitions one block of 256 consecutive rows greedily: a chunk grows while rows <= max_rows,
// entries <= max_nnz and window <= max_window.  Pass 1 (EMIT = false) counts chunks and blob bytes per block, a scan
// turns them into offsets, pass 2 (EMIT = true) repeats the same walk and writes descriptors, the row -> chunk map and
// the 16-bit row offsets inside each blob, plus the compacted list of the chunks that have entries (what the CG kernel walks).  Integer work, ~N x 6 B read + N x 6 B written; negligible next to the solve.
#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kBlockRows = 256;
constexpr int kThreads = 64;

struct Window
{
	uint32_t n;
	uint32_t start[kMaxRanges];
	uint32_t len[kMaxRanges];
	uint32_t total;
};

// window of the rows whose cells lie in [c_a, c_b] (linear keys): for every x[,y] offset one contiguous slot range
// covering cells [c_a + shift - 1, c_b + shift + 1]; overlapping / abutting ranges are merged; ends are made even so that
// every staged segment of doubles is 16-byte aligned.
template<int D>
__device__ __forceinline__ void window_of(const long long c_a, const long long c_b, const uint64_t* __restrict__ cell_start,
	const EnvConst& env, Window& w)
{
	const long long nz = env.grid_n[D - 1];
	const long long ny = (D == 3) ? env.grid_n[1] : 1;
	const long long ncells = static_cast<long long>(env.ncells);
	w.n = 0;
	w.total = 0;
	for (int ox = -1; ox <= 1; ox++)
	{
		for (int oy = (D == 3 ? -1 : 0); oy <= (D == 3 ? 1 : 0); oy++)
		{
			const long long shift = (static_cast<long long>(ox) * ny + oy) * nz;
			long long lo = c_a + shift - 1, hi = c_b + shift + 1;
			if (hi < 0 || lo >= ncells) continue;
			if (lo < 0) lo = 0;
			if (hi > ncells - 1) hi = ncells - 1;
			uint64_t s = cell_start[lo], e = cell_start[hi + 1];
			if (e <= s) continue;
			s &= ~1ull;
			e = (e + 1) & ~1ull;
			if (w.n > 0 && s <= static_cast<uint64_t>(w.start[w.n - 1]) + w.len[w.n - 1])
			{
				const uint64_t prev_end = static_cast<uint64_t>(w.start[w.n - 1]) + w.len[w.n - 1];
				const uint64_t end = e > prev_end ? e : prev_end;
				w.len[w.n - 1] = static_cast<uint32_t>(end - w.start[w.n - 1]);
			}
			else
			{
				w.start[w.n] = static_cast<uint32_t>(s);
				w.len[w.n] = static_cast<uint32_t>(e - s);
				w.n++;
			}
		}
	}
	for (uint32_t k = 0; k < w.n; k++) w.total += w.len[k];
}

struct Open
{
	uint32_t row_begin, rows, nnz;
	uint32_t first_active; // first row with entries (valid when c_a >= 0)
	long long c_a, c_b; // cells of the first / last row that has entries; c_a < 0: none yet
	Window win;
};

template<int D, bool EMIT>
__device__ __forceinline__ void close_chunk(const Open& o, uint32_t& n_chunks, uint32_t& n_live, uint64_t& bytes, uint64_t& cost, const uint64_t chunk_base,
	const uint64_t blob_base, const uint64_t cost_base, const uint64_t live_base, const ChunkLimits& lim,
	const uint32_t* __restrict__ row_len, ChunkDesc* __restrict__ desc, ChunkDesc* __restrict__ live, uint32_t* __restrict__ chunk_of_row,
	unsigned char* __restrict__ blobs)
{
	const uint32_t bb = chunk_blob_bytes(o.rows, o.nnz);
	if (EMIT)
	{
		const uint64_t id = chunk_base + n_chunks;
		ChunkDesc d;
		d.row_begin = o.row_begin; d.rows = o.rows; d.nnz = o.nnz; d.nranges = o.win.n;
		d.blob_off = blob_base + bytes; d.blob_bytes = bb; d.window = o.win.total; d.self_off = 0;
		d.cost_off = cost_base + cost; d.pad_ = 0;
		uint32_t off = 0;
		for (int k = 0; k < kMaxRanges; k++)
		{
			const bool on = static_cast<uint32_t>(k) < o.win.n;
			d.range_start[k] = on ? o.win.start[k] : 0u;
			d.range_len[k] = static_cast<uint16_t>(on ? o.win.len[k] : 0u);
			d.range_off[k] = static_cast<uint16_t>(off);
			// all active rows of the chunk lie in one merged range (their own cells are contiguous slots)
			if (on && o.c_a >= 0 && o.first_active - o.win.start[k] < o.win.len[k])
				d.self_off = static_cast<int32_t>(off + (o.first_active - o.win.start[k])) - static_cast<int32_t>(o.first_active - o.row_begin);
			if (on) off += o.win.len[k];
		}
		desc[id] = d;
		if (o.nnz) live[live_base + n_live] = d;
		*reinterpret_cast<ChunkDesc*>(blobs + d.blob_off) = d; // blob header
		uint16_t* rowoff = reinterpret_cast<uint16_t*>(blobs + d.blob_off + kBlobHeader + static_cast<uint64_t>(round_up8(o.nnz)) * 10u);
		uint32_t run = 0;
		for (uint32_t lr = 0; lr < o.rows; lr++)
		{
			rowoff[lr] = static_cast<uint16_t>(run);
			run += row_len[o.row_begin + lr];
			chunk_of_row[o.row_begin + lr] = static_cast<uint32_t>(id);
		}
		rowoff[o.rows] = static_cast<uint16_t>(run);
	}
	n_chunks += 1;
	bytes += bb;
	if (o.nnz) { n_live += 1; cost += lim.cost_fixed + lim.cost_per_nnz * o.nnz; }
}

template<int D, bool EMIT>
__global__ void __launch_bounds__(kThreads) k_chunk_build(uint64_t n, const uint32_t* __restrict__ row_len, const uint32_t* __restrict__ skey,
	const uint64_t* __restrict__ cell_start, EnvConst env, ChunkLimits lim, uint32_t* __restrict__ blk_chunks, uint32_t* __restrict__ blk_bytes,
	uint32_t* __restrict__ blk_cost, uint32_t* __restrict__ blk_live, const uint64_t* __restrict__ chunk_base, const uint64_t* __restrict__ blob_base,
	const uint64_t* __restrict__ cost_base, const uint64_t* __restrict__ live_base, ChunkDesc* __restrict__ desc, ChunkDesc* __restrict__ live,
	uint64_t desc_cap, uint32_t* __restrict__ chunk_of_row, unsigned char* __restrict__ blobs, DevScalars* sc)
{
	const uint64_t blk = static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x;
	const uint64_t r0 = blk * kBlockRows;
	if (r0 >= n) return;
	const uint64_t r1 = (r0 + kBlockRows < n) ? r0 + kBlockRows : n;
	const uint64_t cbase = EMIT ? chunk_base[blk] : 0, bbase = EMIT ? blob_base[blk] : 0, kbase = EMIT ? cost_base[blk] : 0;
	const uint64_t lbase = EMIT ? live_base[blk] : 0;
	if (EMIT && chunk_base[blk + 1] > desc_cap)
	{
		// cannot happen with the capacity the host allocates unless almost every row needs a chunk of its own
		atomicMax(&sc->error, static_cast<int>(MPS_CUDA_ERROR));
		return;
	}
	uint32_t n_chunks = 0, n_live = 0;
	uint64_t bytes = 0, cost = 0;
	Open o;
	o.row_begin = static_cast<uint32_t>(r0); o.rows = 0; o.nnz = 0; o.c_a = -1; o.c_b = -1; o.win.n = 0; o.win.total = 0; o.first_active = 0;
	for (uint64_t r = r0; r < r1; r++)
	{
		const uint32_t len = row_len[r];
		long long c_a = o.c_a, c_b = o.c_b;
		Window win = o.win;
		uint32_t first_active = o.first_active;
		if (len > 0)
		{
			const long long c = static_cast<long long>(skey[r]); // rows with entries are never Disabled => a real cell
			if (c_a < 0) { c_a = c; first_active = static_cast<uint32_t>(r); }
			if (c != c_b) { c_b = c; window_of<D>(c_a, c_b, cell_start, env, win); }
		}
		const bool fits = (o.rows + 1 <= lim.max_rows) && (o.nnz + len <= lim.max_nnz) && (win.total <= lim.max_window);
		if (!fits && o.rows > 0)
		{
			close_chunk<D, EMIT>(o, n_chunks, n_live, bytes, cost, cbase, bbase, kbase, lbase, lim, row_len, desc, live, chunk_of_row, blobs);
			o.row_begin = static_cast<uint32_t>(r); o.rows = 0; o.nnz = 0; o.c_a = -1; o.c_b = -1; o.win.n = 0; o.win.total = 0;
			c_a = -1; c_b = -1;
			if (len > 0)
			{
				const long long c = static_cast<long long>(skey[r]);
				c_a = c; c_b = c; first_active = static_cast<uint32_t>(r);
				window_of<D>(c_a, c_b, cell_start, env, win);
			}
			else { win.n = 0; win.total = 0; }
			// a single row that does not fit: the host sized the limits from the cell capacity, so this is a logic error
			if (len > lim.max_nnz || win.total > lim.max_window) atomicMax(&sc->error, static_cast<int>(MPS_CUDA_ERROR));
		}
		else if (!fits)
		{
			atomicMax(&sc->error, static_cast<int>(MPS_CUDA_ERROR));
		}
		o.rows += 1; o.nnz += len; o.c_a = c_a; o.c_b = c_b; o.win = win; o.first_active = first_active;
	}
	if (o.rows > 0) close_chunk<D, EMIT>(o, n_chunks, n_live, bytes, cost, cbase, bbase, kbase, lbase, lim, row_len, desc, live, chunk_of_row, blobs);
	if (!EMIT) { blk_chunks[blk] = n_chunks; blk_bytes[blk] = static_cast<uint32_t>(bytes); blk_cost[blk] = static_cast<uint32_t>(cost); blk_live[blk] = n_live; }
}

__global__ void k_chunk_totals(const uint64_t* __restrict__ chunk_base, const uint64_t* __restrict__ blob_base, const uint64_t* __restrict__ cost_base,
	const uint64_t* __restrict__ live_base, uint64_t nblk, DevScalars* sc)
{
	sc->n_chunks = chunk_base[nblk];
	sc->n_live = live_base[nblk];
	sc->blob_total = blob_base[nblk];
	sc->cost_total = cost_base[nblk];
}

#define MPS_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

template<int D>
cudaError_t build(mps_solver* s)
{
	const uint64_t n = s->n;
	CgBuffers& cg = s->cg;
	cudaStream_t st = s->stream;
	const uint64_t nblk = (n + kBlockRows - 1) / kBlockRows;
	const unsigned grid = blocks_for(nblk, kThreads);
	MPS_TRY(cg.blk_chunks.ensure(nblk + 1, st)); MPS_TRY(cg.blk_bytes.ensure(nblk + 1, st));
	MPS_TRY(cg.chunk_base.ensure(nblk + 2, st)); MPS_TRY(cg.blob_base.ensure(nblk + 2, st));
	MPS_TRY(cg.blk_cost.ensure(nblk + 1, st)); MPS_TRY(cg.cost_base.ensure(nblk + 2, st));
	MPS_TRY(cg.blk_live.ensure(nblk + 1, st)); MPS_TRY(cg.live_base.ensure(nblk + 2, st));
	MPS_TRY(cg.chunk_of_row.ensure(n, st));
	// capacities that need no host round trip: entries <= neighbour entries + n; every chunk pads < 16 + 16 + 16 bytes
	const uint64_t desc_cap = n / 16 + nblk + 1024;
	MPS_TRY(cg.desc.ensure(desc_cap, st)); MPS_TRY(cg.live.ensure(desc_cap, st));
	const uint64_t blob_cap = (s->nbr_total + n) * 10 + n * 2 + desc_cap * (96 + kBlobHeader) + 256;
	MPS_TRY(cg.blobs.ensure(blob_cap, st));
	cg.desc_cap = desc_cap;

	k_chunk_build<D, false><<<grid, kThreads, 0, st>>>(n, s->row_len.p, s->skey.p, s->cell_start.p, s->env, cg.limits, cg.blk_chunks.p,
		cg.blk_bytes.p, cg.blk_cost.p, cg.blk_live.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, s->d_sc);
	s->stats.kernel_launches += 1;
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_chunks.p, cg.chunk_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_bytes.p, cg.blob_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_cost.p, cg.cost_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_live.p, cg.live_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	k_chunk_build<D, true><<<grid, kThreads, 0, st>>>(n, s->row_len.p, s->skey.p, s->cell_start.p, s->env, cg.limits, nullptr, nullptr, nullptr, nullptr,
		cg.chunk_base.p, cg.blob_base.p, cg.cost_base.p, cg.live_base.p, cg.desc.p, cg.live.p, desc_cap, cg.chunk_of_row.p, cg.blobs.p, s->d_sc);
	k_chunk_totals<<<1, 1, 0, st>>>(cg.chunk_base.p, cg.blob_base.p, cg.cost_base.p, cg.live_base.p, nblk, s->d_sc);
	s->stats.kernel_launches += 2;
	return cudaGetLastError();
}

} // namespace

cudaError_t launch_chunk_build(mps_solver* s)
{
	if (s->n == 0) return cudaSuccess;
	return s->env.dim == 2 ? build<2>(s) : build<3>(s);
}

} // namespace mps
