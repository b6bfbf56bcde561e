// mps_chunk.cu — cuts the rows of the pressure Poisson equation into "chunks" whose matrix entries AND gathered-vector
// window fit the shared-memory stages of the CG kernel (mps_device.cuh "chunk blob", DESIGN.md "CG kernel").
//
// This replaces what the reference does serially on the host after assembling the rows — inserting them into a uBLAS
// compressed_matrix and copying that into a ViennaCL CSR (Computer.hpp:1337-1354, viennacl/compressed_matrix.hpp:164-191):
// here the "matrix format conversion" is a layout decision made on the device before the rows are written.
//
// Rows are partitioned greedily inside blocks of 256 consecutive rows: a chunk grows while rows <= max_rows,
// entries <= max_nnz and window <= max_window.  ONE WARP per block:
//   1. the lanes stage the block's row lengths and cell keys in shared memory (coalesced), find the distinct cells of the block
//      and gather, per distinct cell and x[,y] column offset, the two cell-table entries a window can start / end at;
//   2. lane 0 walks the rows — the only sequential part, now entirely out of shared memory (round 1 ran this walk with one
//      thread per block straight from global memory: 310 us per launch at 1 M rows, 3 % of the warps active) — and records
//      the chunk boundaries;
//   3. (EMIT) the lanes write descriptors (a lane per chunk), blob headers, the row -> chunk map and the 16-bit row offsets
//      (a lane per row).
// Pass 1 (EMIT = false) counts chunks and blob bytes per block, a scan turns them into offsets, pass 2 (EMIT = true) repeats
// the same walk and writes.  Integer work, ~N x 6 B read + N x 6 B written; negligible next to the solve.
#include "mps_solver.h"

namespace mps {
namespace {

constexpr int kBlockRows = 256;
constexpr int kRowsPerLane = kBlockRows / 32;

struct Window
{
	uint32_t n;
	uint32_t start[kMaxRanges];
	uint32_t len[kMaxRanges];
	uint32_t total;
};

// shared memory of one block (= one warp)
template<int D>
struct BlockSmem
{
	static constexpr int kCols = (D == 3) ? 9 : 3;  // x[,y] column offsets of the stencil
	uint16_t row_len[kBlockRows];
	uint16_t dcell[kBlockRows];                      // row -> index of its distinct cell (rows with entries only)
	uint32_t cell[kBlockRows];                       // distinct cell keys of the block, ascending
	uint32_t cs_lo[kBlockRows][kCols];               // cell_start[clamp(cell + shift - 1)]
	uint32_t cs_hi[kBlockRows][kCols];               // cell_start[clamp(cell + shift + 1) + 1]
	uint32_t row_prefix[kBlockRows + 1];             // exclusive prefix of row_len
	uint16_t lac[kBlockRows];                        // distinct-cell index of the last row with entries at or before this row (0xffff: none)
	uint16_t nar[kBlockRows + 1];                    // first row with entries at or after this row (kBlockRows: none)
	// chunks found by the walk
	uint16_t ch_begin[kBlockRows + 1];               // first row (relative to the block) of chunk q; ch_begin[n_chunks] = rows of the block
	uint16_t ch_da[kBlockRows], ch_db[kBlockRows];   // distinct-cell index of the first / last row with entries (0xffff: none)
	uint16_t ch_first_active[kBlockRows];
	uint16_t ch_live[kBlockRows];                    // rank among the block's chunks that have entries
	uint32_t ch_bytes[kBlockRows];                   // exclusive prefix of blob bytes inside the block
	uint32_t ch_cost[kBlockRows];                    // exclusive prefix of the cost model inside the block
	uint32_t n_chunks, n_live, bytes, cost, error;
};

template<int D> __device__ __forceinline__ long long col_shift(const int k, const EnvConst& env)
{
	const long long nz = env.grid_n[D - 1];
	const long long ny = (D == 3) ? env.grid_n[1] : 1;
	const int ox = (D == 3) ? k / 3 - 1 : k - 1, oy = (D == 3) ? k % 3 - 1 : 0;
	return (static_cast<long long>(ox) * ny + oy) * nz;
}

// window of the rows whose cells lie in [c_a, c_b] (linear keys): for every x[,y] offset one contiguous slot range
// covering cells [c_a + shift - 1, c_b + shift + 1]; overlapping / abutting ranges are merged; ends are made even so that
// every staged segment of doubles is 16-byte aligned.  The cell-table entries come from shared memory (da / db = the
// distinct-cell indices of c_a / c_b).
template<int D>
__device__ __forceinline__ void window_of(const BlockSmem<D>& sm, const uint32_t da, const uint32_t db, const EnvConst& env, Window& w)
{
	const long long ncells = static_cast<long long>(env.ncells);
	const long long c_a = sm.cell[da], c_b = sm.cell[db];
	w.n = 0;
	w.total = 0;
#pragma unroll
	for (int k = 0; k < BlockSmem<D>::kCols; k++)
	{
		const long long shift = col_shift<D>(k, env);
		const long long lo = c_a + shift - 1, hi = c_b + shift + 1;
		if (hi < 0 || lo >= ncells) continue;
		uint64_t s = sm.cs_lo[da][k], e = sm.cs_hi[db][k];
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
	for (uint32_t k = 0; k < w.n; k++) w.total += w.len[k];
}

// the same window, its size only (what the greedy walk needs: registers, no range arrays)
template<int D>
__device__ __forceinline__ uint32_t window_total(const BlockSmem<D>& sm, const uint32_t da, const uint32_t db, const EnvConst& env)
{
	const long long ncells = static_cast<long long>(env.ncells);
	const long long c_a = sm.cell[da], c_b = sm.cell[db];
	uint64_t total = 0, prev_end = 0;
	bool have = false;
#pragma unroll
	for (int k = 0; k < BlockSmem<D>::kCols; k++)
	{
		const long long shift = col_shift<D>(k, env);
		const long long lo = c_a + shift - 1, hi = c_b + shift + 1;
		if (hi < 0 || lo >= ncells) continue;
		uint64_t s = sm.cs_lo[da][k], e = sm.cs_hi[db][k];
		if (e <= s) continue;
		s &= ~1ull;
		e = (e + 1) & ~1ull;
		if (have && s <= prev_end) { if (e > prev_end) { total += e - prev_end; prev_end = e; } }
		else { total += e - s; prev_end = e; have = true; }
	}
	return static_cast<uint32_t>(total);
}

template<int D, bool EMIT>
__global__ void __launch_bounds__(32) k_chunk_build(uint64_t n, const uint32_t* __restrict__ row_len, const uint32_t* __restrict__ skey,
	const uint64_t* __restrict__ cell_start, EnvConst env, ChunkLimits lim, uint32_t* __restrict__ blk_chunks, uint32_t* __restrict__ blk_bytes,
	uint32_t* __restrict__ blk_cost, uint32_t* __restrict__ blk_live, const uint64_t* __restrict__ chunk_base, const uint64_t* __restrict__ blob_base,
	const uint64_t* __restrict__ cost_base, const uint64_t* __restrict__ live_base, ChunkDesc* __restrict__ desc, ChunkDesc* __restrict__ live,
	uint64_t desc_cap, uint32_t* __restrict__ chunk_of_row, unsigned char* __restrict__ blobs, DevScalars* sc, const int skip_empty)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	BlockSmem<D>& sm = *reinterpret_cast<BlockSmem<D>*>(smem_raw);
	constexpr int kCols = BlockSmem<D>::kCols;
	const unsigned lane = threadIdx.x;
	const uint64_t blk = blockIdx.x;
	const uint64_t r0 = blk * kBlockRows;
	if (r0 >= n) return;
	const uint32_t nrows = static_cast<uint32_t>((r0 + kBlockRows < n) ? kBlockRows : n - r0);
	const long long ncells = static_cast<long long>(env.ncells);

	// several ranks: a block without any entry (the other ranks' rows: every rank walks the whole, replicated row range) gets no
	// chunks at all — nobody reads the descriptors or chunk_of_row of rows without entries, and the serial walk below is the cost
	if (skip_empty)
	{
		uint32_t any = 0;
#pragma unroll
		for (int k = 0; k < kRowsPerLane; k++) { const uint32_t lr = lane * kRowsPerLane + k; any |= (lr < nrows) ? row_len[r0 + lr] : 0u; }
		if (!__any_sync(0xffffffffu, any != 0))
		{
			if (!EMIT && lane == 0) { blk_chunks[blk] = 0; blk_bytes[blk] = 0; blk_cost[blk] = 0; blk_live[blk] = 0; }
			return;
		}
	}
	// ---- 1. stage row lengths and cell keys; distinct cells of the rows that have entries; exclusive prefix of the lengths ----
	uint32_t run = 0; // exclusive count of distinct cells / of entries before this lane's rows
	{
		uint32_t len[kRowsPerLane], key[kRowsPerLane];
		uint32_t heads = 0, sum = 0;
		// lane t owns rows [t * 8, t * 8 + 8): keys are non-decreasing along the rows with entries
		uint32_t prev_key = 0xffffffffu;
#pragma unroll
		for (int k = 0; k < kRowsPerLane; k++)
		{
			const uint32_t lr = lane * kRowsPerLane + k;
			len[k] = (lr < nrows) ? row_len[r0 + lr] : 0u;
			key[k] = (lr < nrows && len[k]) ? skey[r0 + lr] : 0xffffffffu;
			sum += len[k];
		}
		// key of the last row with entries before this lane's rows (0xffffffff: none)
		uint32_t last = 0xffffffffu;
#pragma unroll
		for (int k = 0; k < kRowsPerLane; k++) if (key[k] != 0xffffffffu) last = key[k];
		// inclusive "last valid key" scan over the lanes
		uint32_t lastv = last;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t u = __shfl_up_sync(0xffffffffu, lastv, o);
			if (lane >= static_cast<unsigned>(o) && lastv == 0xffffffffu) lastv = u;
		}
		prev_key = __shfl_up_sync(0xffffffffu, lastv, 1);
		if (lane == 0) prev_key = 0xffffffffu;
		uint32_t head_mask = 0;
		{
			uint32_t pk = prev_key;
#pragma unroll
			for (int k = 0; k < kRowsPerLane; k++)
				if (key[k] != 0xffffffffu) { if (key[k] != pk) { head_mask |= 1u << k; heads++; } pk = key[k]; }
		}
		// exclusive scans of heads and sums over the lanes
		uint32_t hinc = heads, sinc = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t hu = __shfl_up_sync(0xffffffffu, hinc, o), su = __shfl_up_sync(0xffffffffu, sinc, o);
			if (lane >= static_cast<unsigned>(o)) { hinc += hu; sinc += su; }
		}
		uint32_t d = hinc - heads; // distinct cells before this lane
		run = sinc - sum;
#pragma unroll
		for (int k = 0; k < kRowsPerLane; k++)
		{
			const uint32_t lr = lane * kRowsPerLane + k;
			if (head_mask & (1u << k)) { sm.cell[d] = key[k]; d++; }
			if (lr < kBlockRows)
			{
				sm.row_len[lr] = static_cast<uint16_t>(len[k]);
				sm.dcell[lr] = (key[k] != 0xffffffffu) ? static_cast<uint16_t>(d - 1) : static_cast<uint16_t>(0xffffu);
				sm.row_prefix[lr] = run;
			}
			run += len[k];
		}
		if (lane == 31) sm.row_prefix[kBlockRows] = run;
		const uint32_t ndist = __shfl_sync(0xffffffffu, hinc, 31);
		__syncwarp();
		// cell-table entries of every distinct cell and column offset
		for (uint32_t q = lane; q < ndist * kCols; q += 32)
		{
			const uint32_t dc = q / kCols; const int k = static_cast<int>(q % kCols);
			const long long c = sm.cell[dc];
			const long long shift = col_shift<D>(k, env);
			long long lo = c + shift - 1, hi = c + shift + 1;
			lo = lo < 0 ? 0 : (lo > ncells - 1 ? ncells - 1 : lo);
			hi = hi < 0 ? 0 : (hi > ncells - 1 ? ncells - 1 : hi);
			sm.cs_lo[dc][k] = static_cast<uint32_t>(cell_start[lo]);
			sm.cs_hi[dc][k] = static_cast<uint32_t>(cell_start[hi + 1]);
		}
	}
	__syncwarp();

	// ---- 2. the greedy partition.  A chunk [begin, end] fits while rows <= max_rows, entries <= max_nnz and window <= max_window;
	//      all three only grow with `end`, so the greedy end is the LAST row that still fits: found by the whole warp in two
	//      rounds (every 8th row, then the 8 rows behind the last hit) instead of a row-by-row walk of one lane (round 1 ran that walk
	//      from global memory: 310 us per launch at 1 M rows; out of shared memory and registers it still was 80 us — the launch is
	//      bound by the instructions of 3 900 single-lane walks).  The sequential part is one iteration per CHUNK (2-6 per block).
	{
		// last cell with entries at or before a row / first row with entries at or after it (lane t owns rows [8 t, 8 t + 8))
		uint16_t lastc = 0xffffu;
		uint32_t firstr = kBlockRows;
		uint16_t dc[kRowsPerLane];
#pragma unroll
		for (int k = 0; k < kRowsPerLane; k++)
		{
			const uint32_t lr = lane * kRowsPerLane + k;
			dc[k] = (lr < nrows) ? sm.dcell[lr] : static_cast<uint16_t>(0xffffu);
			if (dc[k] != 0xffffu) { lastc = dc[k]; if (firstr == kBlockRows) firstr = lr; }
		}
		uint32_t carry_last = lastc, carry_first = firstr;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t ul = __shfl_up_sync(0xffffffffu, carry_last, o), uf = __shfl_down_sync(0xffffffffu, carry_first, o);
			if (lane >= static_cast<unsigned>(o) && carry_last == 0xffffu) carry_last = ul;
			if (lane + o < 32 && carry_first == kBlockRows) carry_first = uf;
		}
		uint32_t before = __shfl_up_sync(0xffffffffu, carry_last, 1); if (lane == 0) before = 0xffffu;
		uint32_t after = __shfl_down_sync(0xffffffffu, carry_first, 1); if (lane == 31) after = kBlockRows;
		uint32_t run_last = before;
#pragma unroll
		for (int k = 0; k < kRowsPerLane; k++)
		{
			if (dc[k] != 0xffffu) run_last = dc[k];
			sm.lac[lane * kRowsPerLane + k] = static_cast<uint16_t>(run_last);
		}
		uint32_t run_first = after;
#pragma unroll
		for (int k = kRowsPerLane - 1; k >= 0; k--)
		{
			if (dc[k] != 0xffffu) run_first = lane * kRowsPerLane + k;
			sm.nar[lane * kRowsPerLane + k] = static_cast<uint16_t>(run_first);
		}
		if (lane == 0) sm.nar[kBlockRows] = static_cast<uint16_t>(kBlockRows);
	}
	__syncwarp();
	{
		uint32_t n_chunks = 0, n_live = 0, bytes = 0, cost = 0, err = 0;
		uint32_t begin = 0;
		// does the chunk [begin, r] fit?  (warp-divergent r, shared memory only)
		auto fits = [&](const uint32_t r) -> bool
		{
			if (r >= nrows) return false;
			const uint32_t rows = r - begin + 1, nnz = sm.row_prefix[r + 1] - sm.row_prefix[begin];
			if (rows > lim.max_rows || nnz > lim.max_nnz) return false;
			const uint32_t fa = sm.nar[begin];
			if (fa > r) return true; // no row with entries: no window
			return window_total<D>(sm, sm.dcell[fa], sm.lac[r], env) <= lim.max_window;
		};
		while (begin < nrows)
		{
			// round 1: rows begin + 8 lane + 7 (lane 31 reaches begin + 255 >= any end); the hits are a prefix of the lanes
			const unsigned hit8 = __ballot_sync(0xffffffffu, fits(begin + lane * 8u + 7u));
			const uint32_t base = begin + 8u * static_cast<uint32_t>(__popc(hit8)); // rows [begin, base) fit as a whole (or base == begin)
			// round 2: rows base .. base + 7
			const unsigned hit1 = __ballot_sync(0xffffffffu, lane < 8 && fits(base + lane));
			uint32_t end = base + static_cast<uint32_t>(__popc(hit1)); // one past the last row of the chunk
			if (end == begin)
			{
				// a single row that does not fit: the host sized the limits from the cell capacity, so this is a logic error
				err = 1; end = begin + 1;
			}
			const uint32_t rows = end - begin, nnz = sm.row_prefix[end] - sm.row_prefix[begin];
			const uint32_t fa = sm.nar[begin];
			const bool active = fa < end;
			if (lane == 0)
			{
				sm.ch_begin[n_chunks] = static_cast<uint16_t>(begin);
				sm.ch_da[n_chunks] = active ? sm.dcell[fa] : static_cast<uint16_t>(0xffffu);
				sm.ch_db[n_chunks] = active ? sm.lac[end - 1] : static_cast<uint16_t>(0xffffu);
				sm.ch_first_active[n_chunks] = static_cast<uint16_t>(active ? fa : begin);
				sm.ch_live[n_chunks] = static_cast<uint16_t>(n_live);
				sm.ch_bytes[n_chunks] = bytes; sm.ch_cost[n_chunks] = cost;
			}
			n_chunks++;
			bytes += chunk_blob_bytes(rows, nnz);
			if (nnz) { n_live++; cost += lim.cost_fixed + lim.cost_per_nnz * nnz; }
			begin = end;
		}
		if (lane == 0)
		{
			sm.ch_begin[n_chunks] = static_cast<uint16_t>(nrows);
			sm.n_chunks = n_chunks; sm.n_live = n_live; sm.bytes = bytes; sm.cost = cost; sm.error = err;
		}
	}
	__syncwarp();
	if (sm.error && lane == 0) atomicMax(&sc->error, static_cast<int>(MPS_CUDA_ERROR));
	if (!EMIT)
	{
		if (lane == 0) { blk_chunks[blk] = sm.n_chunks; blk_bytes[blk] = sm.bytes; blk_cost[blk] = sm.cost; blk_live[blk] = sm.n_live; }
		return;
	}

	// ---- 3. emit ----
	const uint64_t cbase = chunk_base[blk], bbase = blob_base[blk], kbase = cost_base[blk], lbase = live_base[blk];
	if (chunk_base[blk + 1] > desc_cap)
	{
		// cannot happen with the capacity the host allocates unless almost every row needs a chunk of its own
		if (lane == 0) atomicMax(&sc->error, static_cast<int>(MPS_CUDA_ERROR));
		return;
	}
	if (sm.error) return; // an oversized chunk must never reach the CG kernel's fixed-size stages
	const uint32_t n_chunks = sm.n_chunks;
	for (uint32_t q = lane; q < n_chunks; q += 32)
	{
		const uint32_t begin = sm.ch_begin[q], rows = sm.ch_begin[q + 1] - begin;
		const uint32_t nnz = sm.row_prefix[begin + rows] - sm.row_prefix[begin];
		Window w; w.n = 0; w.total = 0;
		const bool active = sm.ch_da[q] != 0xffffu;
		if (active) window_of<D>(sm, sm.ch_da[q], sm.ch_db[q], env, w);
		const uint32_t first_active = static_cast<uint32_t>(r0) + sm.ch_first_active[q];
		ChunkDesc d;
		d.row_begin = static_cast<uint32_t>(r0) + begin; d.rows = rows; d.nnz = nnz; d.nranges = w.n;
		d.blob_off = bbase + sm.ch_bytes[q]; d.blob_bytes = chunk_blob_bytes(rows, nnz); d.window = w.total; d.self_off = 0;
		d.cost_off = kbase + sm.ch_cost[q]; d.pad_ = 0;
		uint32_t off = 0;
		for (int k = 0; k < kMaxRanges; k++)
		{
			const bool on = static_cast<uint32_t>(k) < w.n;
			d.range_start[k] = on ? w.start[k] : 0u;
			d.range_len[k] = static_cast<uint16_t>(on ? w.len[k] : 0u);
			d.range_off[k] = static_cast<uint16_t>(off);
			// all active rows of the chunk lie in one merged range (their own cells are contiguous slots)
			if (on && active && first_active - w.start[k] < w.len[k])
				d.self_off = static_cast<int32_t>(off + (first_active - w.start[k])) - static_cast<int32_t>(first_active - d.row_begin);
			if (on) off += w.len[k];
		}
		desc[cbase + q] = d;
		if (nnz) live[lbase + sm.ch_live[q]] = d;
		*reinterpret_cast<ChunkDesc*>(blobs + d.blob_off) = d; // blob header
		uint16_t* rowoff = reinterpret_cast<uint16_t*>(blobs + d.blob_off + kBlobHeader + static_cast<uint64_t>(round_up8(nnz)) * 10u);
		rowoff[rows] = static_cast<uint16_t>(nnz);
	}
	// per row: its chunk and its 16-bit offset inside the chunk's blob (a lane per row; the chunk of a row by binary search)
	for (uint32_t lr = lane; lr < nrows; lr += 32)
	{
		uint32_t lo = 0, hi = n_chunks; // last chunk with ch_begin <= lr
		while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sm.ch_begin[mid] <= lr) lo = mid; else hi = mid; }
		const uint32_t begin = sm.ch_begin[lo], rows = sm.ch_begin[lo + 1] - begin;
		const uint32_t nnz = sm.row_prefix[begin + rows] - sm.row_prefix[begin];
		chunk_of_row[r0 + lr] = static_cast<uint32_t>(cbase + lo);
		uint16_t* rowoff = reinterpret_cast<uint16_t*>(blobs + bbase + sm.ch_bytes[lo] + kBlobHeader + static_cast<uint64_t>(round_up8(nnz)) * 10u);
		rowoff[lr - begin] = static_cast<uint16_t>(sm.row_prefix[lr] - sm.row_prefix[begin]);
	}
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
	const unsigned grid = static_cast<unsigned>(nblk); // one warp per block of rows
	const size_t smem = sizeof(BlockSmem<D>);
	static bool attr_set[2] = { false, false };
	if (!attr_set[D - 2])
	{
		MPS_TRY(cudaFuncSetAttribute(k_chunk_build<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		MPS_TRY(cudaFuncSetAttribute(k_chunk_build<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
		attr_set[D - 2] = true;
	}
	MPS_TRY(cg.blk_chunks.ensure(nblk + 1, st)); MPS_TRY(cg.blk_bytes.ensure(nblk + 1, st));
	MPS_TRY(cg.chunk_base.ensure(nblk + 2, st)); MPS_TRY(cg.blob_base.ensure(nblk + 2, st));
	MPS_TRY(cg.blk_cost.ensure(nblk + 1, st)); MPS_TRY(cg.cost_base.ensure(nblk + 2, st));
	MPS_TRY(cg.blk_live.ensure(nblk + 1, st)); MPS_TRY(cg.live_base.ensure(nblk + 2, st));
	MPS_TRY(cg.chunk_of_row.ensure(n, st));
	// capacities that need no host round trip: entries <= neighbour entries + n; every chunk pads < 16 + 16 + 16 bytes
	const uint64_t desc_cap = n / 16 + nblk + 1024;
	MPS_TRY(cg.desc.ensure(desc_cap, st)); MPS_TRY(cg.live.ensure(desc_cap, st));
	uint64_t entries = s->nbr_total + (s->own1() - s->own0()); // bound: the list also holds the candidates between r_e and the cell size
	if (alloc_slack_percent() == 0)
	{
		// blocks that only just fit the GPU (MPS_ALLOC_SLACK=0): size the blobs from the exact entry count (the row-pointer scan has
		// just produced it) at the price of one more host read-back per step — in 3-D the bound is 1.6x the matrix
		uint64_t nnz = 0;
		MPS_TRY(cudaMemcpyAsync(&nnz, cg.rowptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		MPS_TRY(cudaStreamSynchronize(st));
		if (nnz < entries) entries = nnz;
	}
	const uint64_t blob_cap = entries * 10 + n * 2 + desc_cap * (96 + kBlobHeader) + 256;
	MPS_TRY(cg.blobs.ensure(blob_cap, st));
	cg.desc_cap = desc_cap;

	k_chunk_build<D, false><<<grid, 32, smem, st>>>(n, s->row_len.p, s->skey.p, s->cell_start.p, s->env, cg.limits, cg.blk_chunks.p,
		cg.blk_bytes.p, cg.blk_cost.p, cg.blk_live.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, s->d_sc, s->comm.on ? 1 : 0);
	s->stats.kernel_launches += 1;
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_chunks.p, cg.chunk_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_bytes.p, cg.blob_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_cost.p, cg.cost_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	MPS_TRY(launch_exclusive_scan_u32_to_u64(cg.blk_live.p, cg.live_base.p, nblk, s->scan_tmp, st, &s->stats.kernel_launches));
	k_chunk_build<D, true><<<grid, 32, smem, st>>>(n, s->row_len.p, s->skey.p, s->cell_start.p, s->env, cg.limits, nullptr, nullptr, nullptr, nullptr,
		cg.chunk_base.p, cg.blob_base.p, cg.cost_base.p, cg.live_base.p, cg.desc.p, cg.live.p, desc_cap, cg.chunk_of_row.p, cg.blobs.p, s->d_sc, s->comm.on ? 1 : 0);
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
