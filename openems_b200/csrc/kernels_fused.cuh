// kernels_fused.cuh -- one-pass timestep: E half-step and H half-step fused into one kernel.
//
// Traffic per cell and timestep drops from 76 B (two passes: each reads E and H and writes one
// of them) to 50 B (read E, H and the operator index once, write E and H once).  The arithmetic
// per cell is exactly that of k_update_E / k_update_H (same helpers, same roundings), only the
// order in which cells are visited changes, so the results are bit-identical.
//
// Scheme (z-marching, out of place: the field sets are ping-pong buffers, the source set is
// never written during a step):
//   block = (32 lanes, TY+1 warps), tile = 128 x cells by TY rows, marching a z chunk upwards.
//   iteration kk:  every warp computes E_new(kk) of its row (4 cells per lane) from E_old(kk),
//                  H_old(kk), H_old(kk-1) [registers], stores it (rows of the tile) and publishes
//                  V0/V2 in shared memory; warp TY does the same for the halo row j0+TY without
//                  storing; lane 31 also computes V1/V2 of the halo column cell x0+128.
//                  __syncthreads
//                  H_new(kk-1) of the tile from E_new(kk-1) [registers], E_new(kk) [registers],
//                  E_new(kk-1) of row j+1 [shared memory], of cell i+1 [shuffle / halo column].
// Hooks that touch E between the two half-steps (Mur, excitation) run after this kernel; the H
// cells that depend on an E value they changed are recomputed by k_fix_H from the untouched
// source set (engine.cu builds that list at upload).  Lorentz/RLC switch the engine back to
// the two-pass schedule.
//
// UPML (4.6 % of the cells of the 1024^3 PML_8 benchmark) stays on a two-pass "shell" around
// the one-pass interior, so that the big kernel carries no UPML code at all:
//   1. k_shell_E  all UPML boxes: E_new of the box cells IN PLACE in the source set (E_old of a cell
//                                is read by nobody but the cell itself), flux in place
//   2. k_fused_EH              : all other cells; shell cells are passed through (their E is
//                                already E_new and is copied to the destination set, their H is
//                                copied unchanged)
//   3. (hooks, k_fix_H)
//   4. k_shell_H  all UPML boxes: H_new of the box cells from the FINAL E of the destination set,
//                                source set -> destination set
#pragma once
#include "kernels.cuh"

#ifndef FUSED_TY
#define FUSED_TY 7   // rows per tile; +1 halo-row warp = 8 warps = 256 threads, so 2 blocks/SM may use 128 registers
#endif

struct FusedParams {
	const float* Vs; const float* Is;   // source set (timestep n)
	float* Vd; float* Id;               // destination set (timestep n+1)
	const void* idx;
	const float4 *eA, *eB;              // E tables (vv|UPML flag, vi)
	const float4 *hA, *hB;              // H tables (ii|UPML flag, iv)
	int nx, ny, nz;      // nz = local planes held
	int pitch;
	long long plane, comp;
	int kE0, kE1;        // local planes whose E this kernel produces (owned planes)
	int kH0, kH1;        // local planes whose H is UPDATED here; other owned planes are copied through
	int kHc1;            // H planes [kH1, kHc1) are copied through (top of the domain); [kHc1, kE1) left to the slab kernel
	int zchunk;
	// rows whose E and H this kernel stores: [jb, je) (default 0, ny).  UPML boxes that span whole rows / planes at
	// the ends of the mesh are skipped altogether (the host also narrows kE0 / kE1): k_shell_E stores their E_new in
	// both field sets, k_shell_H their H_new, so nothing of them has to be passed through here
	int jb, je;
	// UPML shell: regions (whole float4 chunks in x) whose E and H are produced by k_shell_E/H
	int nsh;
	struct ShellBox { int c0, cn, j0, jn, k0, kn; } sh[OEMS_MAX_PML_BOXES];
};

// out-of-place UPML update of one component: takes the flux of timestep n, returns the new field
// value and the new flux (engine_ext_upml.cpp:52-137 / :144-229 around the leapfrog)
__device__ __forceinline__ float leap_pml_oop(float X, float m_vv, float m_vi, float curl, float a_vv, float a_fn,
                                              float a_fo, float F, float& Fn)
{
	const float f = fsub(fmul(a_vv, X), fmul(a_fo, F));
	Fn = fadd(fmul(F, m_vv), fmul(m_vi, curl));
	return fadd(f, fmul(a_fn, Fn));
}

template <typename IdxT, bool HAS_PML>
__global__ void __launch_bounds__(32 * (FUSED_TY + 1), 2) k_fused_EH(const __grid_constant__ FusedParams p)
{
	PDL_PROLOGUE();
	__shared__ float4 xV0[3][FUSED_TY + 1][32];
	__shared__ float4 xV2[3][FUSED_TY + 1][32];

	const int lane = threadIdx.x, ty = threadIdx.y;
	const int i0 = (blockIdx.x * 32 + lane) * 4;
	const int j = blockIdx.y * FUSED_TY + ty;
	const bool halo_row = ty == FUSED_TY;
	const int kb = p.kE0 + blockIdx.z * p.zchunk;
	const int ke = min(kb + p.zchunk, p.kE1);
	if (kb >= ke) return; // block-uniform
	// H planes of this chunk that are updated / copied through
	const int he = min(ke, p.kH1);
	// E planes to compute: the chunk's own planes, plus plane ke when H(ke-1) is updated here
	const int e_last = (he == ke && ke < p.nz) ? ke : ke - 1;

	const bool row_ok = j < p.ny;
	const bool active = row_ok && i0 < p.pitch;
	const int ic = i0 < p.pitch ? i0 : 0;
	const int jc = row_ok ? j : p.ny - 1;
	const int jm = jc - (jc > 0);
	const long long row = (long long)jc * p.pitch + ic;
	const long long rowm = (long long)jm * p.pitch + ic;
	const float* __restrict__ I0 = p.Is;
	const float* __restrict__ I1 = p.Is + p.comp;
	const float* __restrict__ I2 = p.Is + 2 * p.comp;
	const float* __restrict__ V0 = p.Vs;
	const float* __restrict__ V1 = p.Vs + p.comp;
	const float* __restrict__ V2 = p.Vs + 2 * p.comp;
	// halo column: cell xe = i0 + 4 of lane 31 (first cell of the next tile in x)
	const int xe = ic + 4;
	const bool hcol = lane == 31 && !halo_row && active && xe < p.nx;

	// shell membership of this thread's chunk (shb) and of the chunk to its right (shxb), one bit
	// per plane of the z chunk (zchunk + 1 <= 64 planes, checked by the host)
	unsigned long long shb = 0, shxb = 0;
	if (HAS_PML) {
		const int chunk = ic >> 2;
		for (int b = 0; b < p.nsh; ++b) {
			if ((unsigned)(jc - p.sh[b].j0) >= (unsigned)p.sh[b].jn) continue;
			const bool mine = (unsigned)(chunk - p.sh[b].c0) < (unsigned)p.sh[b].cn;
			const bool right = (unsigned)(chunk + 1 - p.sh[b].c0) < (unsigned)p.sh[b].cn;
			if (!mine && !right) continue;
			const int a = max(p.sh[b].k0, kb) - kb, z = min(p.sh[b].k0 + p.sh[b].kn, e_last + 1) - kb;
			if (z <= a) continue;
			const unsigned long long m = (z - a >= 64 ? ~0ull : ((1ull << (z - a)) - 1ull)) << a;
			if (mine) shb |= m;
			if (right) shxb |= m;
		}
	}
	bool shk = false; // plane k (previous iteration) of this chunk belongs to the shell

	// carried state: E_new(k) and H_old(k) of the own cells, operator indices of plane k
	float4 ek0 = make_float4(0, 0, 0, 0), ek1 = ek0, ek2 = ek0;
	float4 hk0, hk1, hk2 = make_float4(0, 0, 0, 0);
	unsigned ek_idx[4] = {0, 0, 0, 0};
	float hcV1 = 0.0f, hcV2 = 0.0f, hcI0 = 0.0f;
	{
		const int km = kb - (kb > 0); // z-1 clamp only at the bottom of the (local) domain
		const long long o = (long long)km * p.plane + row;
		hk0 = ld4(I0 + o);
		hk1 = ld4(I1 + o);
		if (hcol) hcI0 = I0[o + 4];
	}

	for (int kk = kb; kk <= e_last; ++kk) {
		// ------------------------------------------------------------ E_new(kk)
		const long long o = (long long)kk * p.plane + row;
		const long long om = (long long)kk * p.plane + rowm;
		// shell cells: E_new was stored by k_shell_E before this kernel started
		const bool sh = HAS_PML && (shb >> (kk - kb) & 1ull), shx = HAS_PML && (shxb >> (kk - kb) & 1ull);
		unsigned e[4];
		Idx4<IdxT>::load(p.idx, o, e);
		if (kk + 1 <= e_last && (lane & 7) == 0) {
			const long long of = o + p.plane;
			prefetch_l2(I0 + of); prefetch_l2(I1 + of); prefetch_l2(I2 + of);
			prefetch_l2(V0 + of); prefetch_l2(V1 + of); prefetch_l2(V2 + of);
			if ((lane & 15) == 0) prefetch_l2(reinterpret_cast<const IdxT*>(p.idx) + of);
		}
		const float4 i0c = ld4(I0 + o), i1c = ld4(I1 + o), i2c = ld4(I2 + o);
		const float4 i0jm = ld4(I0 + om), i2jm = ld4(I2 + om);
		float4 v0 = ld4(V0 + o), v1 = ld4(V1 + o), v2 = ld4(V2 + o); // shell cells: already E_new (k_shell_E works in place)
		float l1 = __shfl_up_sync(0xffffffffu, i1c.w, 1);
		float l2 = __shfl_up_sync(0xffffffffu, i2c.w, 1);
		if (lane == 0) {
			if (ic > 0) { l1 = I1[o - 1]; l2 = I2[o - 1]; }
			else { l1 = i1c.x; l2 = i2c.x; }
		}
		const float4 i1xm = make_float4(l1, i1c.x, i1c.y, i1c.z);
		const float4 i2xm = make_float4(l2, i2c.x, i2c.y, i2c.z);
		float nV1 = 0.0f, nV2 = 0.0f, nI0 = 0.0f; // halo column values of plane kk
		if (active) {
			const bool uni = (e[0] == e[1]) & (e[1] == e[2]) & (e[2] == e[3]);
			float4 A = __ldg(p.eA + e[0]), B = __ldg(p.eB + e[0]);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				if (c > 0 && !uni) { A = __ldg(p.eA + e[c]); B = __ldg(p.eB + e[c]); }
				const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(hk1, c));
				const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(hk0, c)), comp(i2c, c)), comp(i2xm, c));
				const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
				// branch-free: the shell keeps what it loaded
				const float n0 = leap(comp(v0, c), A.x, B.x, curl0), n1 = leap(comp(v1, c), A.y, B.y, curl1), n2 = leap(comp(v2, c), A.z, B.z, curl2);
				setcomp(v0, c, sh ? comp(v0, c) : n0);
				setcomp(v1, c, sh ? comp(v1, c) : n1);
				setcomp(v2, c, sh ? comp(v2, c) : n2);
			}
			if (!halo_row && kk < ke) {
				st4(p.Vd + o, v0);
				st4(p.Vd + p.comp + o, v1);
				st4(p.Vd + 2 * p.comp + o, v2);
			}
			if (hcol) {
				// V1, V2 of cell (xe, j, kk): engine.cpp:148-166 with the x-1 neighbour = my last cell
				const unsigned ex = reinterpret_cast<const IdxT*>(p.idx)[o + 4];
				const float xi0 = I0[o + 4], xi1 = I1[o + 4], xi2 = I2[o + 4], xi0jm = I0[om + 4]; // V0 of the halo cell is not needed
				const float c1x = fadd(fsub(fsub(xi0, hcI0), xi2), i2c.w);
				const float c2x = fadd(fsub(fsub(xi1, i1c.w), xi0), xi0jm);
				const float4 Ax = __ldg(p.eA + ex), Bx = __ldg(p.eB + ex);
				const float b1 = V1[o + 4], b2 = V2[o + 4];
				const float n1 = leap(b1, Ax.y, Bx.y, c1x), n2 = leap(b2, Ax.z, Bx.z, c2x);
				nV1 = shx ? b1 : n1;
				nV2 = shx ? b2 : n2;
				nI0 = xi0;
			}
		}
		xV0[kk % 3][ty][lane] = v0;
		xV2[kk % 3][ty][lane] = v2;
		__syncthreads();

		// ------------------------------------------------------------ H_new(kk-1)
		const int k = kk - 1;
		// i+1 neighbours of E_new(k): next lane's first cell (whole warp takes part in the shuffle)
		float r1 = __shfl_down_sync(0xffffffffu, ek1.x, 1);
		float r2 = __shfl_down_sync(0xffffffffu, ek2.x, 1);
		if (lane == 31) { r1 = hcV1; r2 = hcV2; }
		if (k >= kb && !halo_row && active) {
			const long long oh = (long long)k * p.plane + row;
			float4 c0 = hk0, c1 = hk1, c2 = hk2;
			if (k < he && j < p.ny - 1 && !shk) { // H of shell cells: copied through here, updated by k_shell_H
				const float4 v0jp = xV0[k % 3][ty + 1][lane], v2jp = xV2[k % 3][ty + 1][lane];
				const float4 v1xp = make_float4(ek1.y, ek1.z, ek1.w, r1);
				const float4 v2xp = make_float4(ek2.y, ek2.z, ek2.w, r2);
				const bool uni = (ek_idx[0] == ek_idx[1]) & (ek_idx[1] == ek_idx[2]) & (ek_idx[2] == ek_idx[3]);
				float4 A = __ldg(p.hA + ek_idx[0]), B = __ldg(p.hB + ek_idx[0]);
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					if (c > 0 && !uni) { A = __ldg(p.hA + ek_idx[c]); B = __ldg(p.hB + ek_idx[c]); }
					if (ic + c < p.nx - 1) {
						const float curl0 = fadd(fsub(fsub(comp(ek2, c), comp(v2jp, c)), comp(ek1, c)), comp(v1, c));
						const float curl1 = fadd(fsub(fsub(comp(ek0, c), comp(v0, c)), comp(ek2, c)), comp(v2xp, c));
						const float curl2 = fadd(fsub(fsub(comp(ek1, c), comp(v1xp, c)), comp(ek0, c)), comp(v0jp, c));
						setcomp(c0, c, leap(comp(c0, c), A.x, B.x, curl0));
						setcomp(c1, c, leap(comp(c1, c), A.y, B.y, curl1));
						setcomp(c2, c, leap(comp(c2, c), A.z, B.z, curl2));
					}
				}
			}
			if (k < p.kHc1) {
				st4(p.Id + oh, c0);
				st4(p.Id + p.comp + oh, c1);
				st4(p.Id + 2 * p.comp + oh, c2);
			}
		}
		// rotate: plane kk becomes "k"
		ek0 = v0; ek1 = v1; ek2 = v2;
		hk0 = i0c; hk1 = i1c; hk2 = i2c;
#pragma unroll
		for (int c = 0; c < 4; ++c) ek_idx[c] = e[c];
		hcV1 = nV1; hcV2 = nV2; hcI0 = nI0;
		shk = sh;
	}
	// H of the chunk's last plane when it is not updated in this kernel (top of the domain:
	// copy through; top plane of a slab with an upper neighbour: left to the slab kernel)
	const int k = e_last;
	if (k == ke - 1 && k >= kb && !halo_row && active && k < p.kHc1) {
		const long long oh = (long long)k * p.plane + row;
		st4(p.Id + oh, hk0);
		st4(p.Id + p.comp + oh, hk1);
		st4(p.Id + 2 * p.comp + oh, hk2);
	}
}

// ---------------------------------------------------------------------------------------
// H of listed cells recomputed out of place from the source set and the FINAL E of the
// destination set: the cells whose E neighbours were changed by hooks (Mur, excitation) after
// k_fused_EH ran.  One thread per listed cell, all three components.  UPML cells are skipped:
// k_shell_H runs after the hooks and sees the final E anyway.
// ---------------------------------------------------------------------------------------
struct FixParams {
	const float* Is; float* Id;
	const float* Vd;                 // final E of this timestep
	const void* idx;
	const float4 *hA, *hB;
	const int* cell;                 // [count][3] x, y, local z
	long long count;
	int pitch; long long plane, comp;
};

template <typename IdxT>
__global__ void k_fix_H(const __grid_constant__ FixParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	const int x = p.cell[3 * t], j = p.cell[3 * t + 1], k = p.cell[3 * t + 2];
	const long long o = (long long)k * p.plane + (long long)j * p.pitch + x;
	const float* V0 = p.Vd; const float* V1 = p.Vd + p.comp; const float* V2 = p.Vd + 2 * p.comp;
	const unsigned e = reinterpret_cast<const IdxT*>(p.idx)[o];
	const float4 A = __ldg(p.hA + e), B = __ldg(p.hB + e);
	if (A.w != 0.0f) return;
	const float v0 = V0[o], v1 = V1[o], v2 = V2[o];
	const float curl0 = fadd(fsub(fsub(v2, V2[o + p.pitch]), v1), V1[o + p.plane]);
	const float curl1 = fadd(fsub(fsub(v0, V0[o + p.plane]), v2), V2[o + 1]);
	const float curl2 = fadd(fsub(fsub(v1, V1[o + 1]), v0), V0[o + p.pitch]);
	p.Id[o] = leap(p.Is[o], A.x, B.x, curl0);
	p.Id[p.comp + o] = leap(p.Is[p.comp + o], A.y, B.y, curl1);
	p.Id[2 * p.comp + o] = leap(p.Is[2 * p.comp + o], A.z, B.z, curl2);
}

// ---------------------------------------------------------------------------------------
// UPML shell of the one-pass schedule: the two half-steps of ONE UPML box, out of place
// (field of the source set -> destination set), flux updated in place (every cell owns its
// flux values).  Same arithmetic as the UPML branch of k_update_E / k_update_H.
// Thread mapping: XL lanes side by side in x (4 cells each), 32/XL rows per warp, so that the
// 8-cell-thin x slabs still fill their warps; z-marching with the k-1 / k+1 plane in registers.
// The launch covers the whole float4 chunks the box touches in x: cells of those chunks that
// lie in no UPML box get the plain leapfrog here (k_fused_EH skips whole chunks), cells of
// another box are left to that box's launch.
// ---------------------------------------------------------------------------------------
#ifndef SHELL_MIN_BLOCKS
#define SHELL_MIN_BLOCKS 3
#endif
#define OEMS_MAX_SHELL_ENTRIES 16   // boxes + the pieces of x-window boxes that lie in skipped planes / rows
struct ShellBoxParams {
	float* flux;       // component 0 of this box
	float* flux_out;   // k_shell_E: where the new flux goes (= flux, except for the pieces of x-window boxes, whose voltage flux is ping-ponged with the field sets)
	long long cs;      // flux component stride = cells of the box held here
	int bs0, bs1, bs2; // box origin: x, y, local z (origin of the flux layout [k][j][i])
	int bn0, bn1;      // box lines in x, y
	int jr0, jr1;      // rows of the box to process (box-local), normally 0 .. bn1
	int k0, k1;        // local planes to process
	int c0, nchunk;    // float4 chunks that cover the box in x
	int zchunk;
	int xl;            // lanes side by side in x (power of two)
	int gx, gy;        // blocks in x and y; blocks of this box = gx*gy*gz
	unsigned blk0;     // first block of this box in the launch
};
struct ShellParams {
	const float* Xs;   // field being updated, source set (E step: E_old, H step: H_old)
	float* Xd;         // destination set
	const float* Y;    // the other field (E step: H_old, H step: final E_new)
	const void* idx;
	const float4 *tA, *tB, *tP0, *tP1, *tP2;
	int nx, ny, pitch;
	long long plane, comp;
	int nboxes;
	unsigned nblocks;
	// k_shell_E: cells outside the planes [sk0, sk1) / rows [sjb, sje) are skipped by the one-pass kernels: their E_new
	// is stored in the destination set Xd2 (NULL: nothing is skipped) instead of in place -- in both sets on the
	// plane sk1 / row sje, which the one-pass kernels read from the source set as +1 neighbours
	float* Xd2;
	int sk0, sk1, sjb, sje;
	ShellBoxParams box[OEMS_MAX_SHELL_ENTRIES];
};

// block -> (box, block coordinates inside the box)
struct ShellBlock { int b, bx, by, bz; };
__device__ __forceinline__ ShellBlock shell_block(const ShellParams& p)
{
	ShellBlock r;
	r.b = 0;
	while (r.b + 1 < p.nboxes && blockIdx.x >= p.box[r.b + 1].blk0) ++r.b;
	unsigned l = blockIdx.x - p.box[r.b].blk0;
	r.bx = (int)(l % (unsigned)p.box[r.b].gx); l /= (unsigned)p.box[r.b].gx;
	r.by = (int)(l % (unsigned)p.box[r.b].gy);
	r.bz = (int)(l / (unsigned)p.box[r.b].gy);
	return r;
}

// true if a box before b covers (chunk, j, k) with its chunk-aligned footprint
__device__ __forceinline__ bool shell_earlier_box(const ShellParams& p, int b, int chunk, int j, int k)
{
	for (int a = 0; a < b; ++a) {
		const ShellBoxParams& q = p.box[a];
		if ((unsigned)(chunk - q.c0) < (unsigned)q.nchunk && (unsigned)(j - q.bs1) < (unsigned)q.bn1 && k >= q.k0 && k < q.k1) return true;
	}
	return false;
}

template <typename IdxT>
__global__ void __launch_bounds__(256, SHELL_MIN_BLOCKS) k_shell_E(const __grid_constant__ ShellParams p)
{
	PDL_PROLOGUE();
	const ShellBlock sb = shell_block(p);
	const ShellBoxParams& q = p.box[sb.b];
	const int XL = q.xl;
	const int lane = threadIdx.x, sub = lane & (XL - 1);
	const int ch = sb.bx * XL + sub;
	const int lj = q.jr0 + (sb.by * blockDim.y + threadIdx.y) * (32 / XL) + lane / XL;
	const int kb = q.k0 + sb.bz * q.zchunk;
	const int ke = min(kb + q.zchunk, q.k1);
	if (kb >= ke) return;
	const bool active = ch < q.nchunk && lj < q.jr1;
	if (__all_sync(0xffffffffu, !active)) return;
	const int ic = (q.c0 + (ch < q.nchunk ? ch : 0)) * 4;
	const int j = q.bs1 + (lj < q.jr1 ? lj : q.jr0);
	const int jm = j - (j > 0);
	const long long row = (long long)j * p.pitch + ic;
	const long long rowm = (long long)jm * p.pitch + ic;
	const float* __restrict__ I0 = p.Y;
	const float* __restrict__ I1 = p.Y + p.comp;
	const float* __restrict__ I2 = p.Y + 2 * p.comp;
	const float* __restrict__ V0 = p.Xs;
	const float* __restrict__ V1 = p.Xs + p.comp;
	const float* __restrict__ V2 = p.Xs + 2 * p.comp;

	float4 i0km, i1km;
	{
		const int km = kb - (kb > 0);
		const long long o = (long long)km * p.plane + row;
		i0km = ld4(I0 + o);
		i1km = ld4(I1 + o);
	}
	for (int k = kb; k < ke; ++k) {
		const long long o = (long long)k * p.plane + row;
		const long long om = (long long)k * p.plane + rowm;
		unsigned e[4];
		Idx4<IdxT>::load(p.idx, o, e);
		if (XL >= 16 && k + 1 < ke && active && (sub & 7) == 0) { // thin boxes: a line prefetch would fetch 128 B for the 48 B they use
			const long long of = o + p.plane;
			prefetch_l2(I0 + of); prefetch_l2(I1 + of); prefetch_l2(I2 + of);
			prefetch_l2(V0 + of); prefetch_l2(V1 + of); prefetch_l2(V2 + of);
			prefetch_l2(reinterpret_cast<const IdxT*>(p.idx) + of);
			const float* fn = q.flux + ((long long)(k + 1 - q.bs2) * q.bn1 + lj) * q.bn0 + max(ic - q.bs0, 0);
			prefetch_l2(fn); prefetch_l2(fn + q.cs); prefetch_l2(fn + 2 * q.cs);
		}
		const float4 i0c = ld4(I0 + o), i1c = ld4(I1 + o), i2c = ld4(I2 + o);
		const float4 i0jm = ld4(I0 + om), i2jm = ld4(I2 + om);
		float4 v0 = ld4(V0 + o), v1 = ld4(V1 + o), v2 = ld4(V2 + o);
		float l1 = __shfl_up_sync(0xffffffffu, i1c.w, 1, XL);
		float l2 = __shfl_up_sync(0xffffffffu, i2c.w, 1, XL);
		if (sub == 0) {
			if (ic > 0) { l1 = I1[o - 1]; l2 = I2[o - 1]; }
			else { l1 = i1c.x; l2 = i2c.x; }
		}
		const float4 i1xm = make_float4(l1, i1c.x, i1c.y, i1c.z);
		const float4 i2xm = make_float4(l2, i2c.x, i2c.y, i2c.z);
		if (active) {
			const long long frow = ((long long)(k - q.bs2) * q.bn1 + lj) * q.bn0;
			unsigned done = 0;
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const int li = ic + c - q.bs0;
				const float4 A = __ldg(p.tA + e[c]), B = __ldg(p.tB + e[c]);
				const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(i1km, c));
				const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(i0km, c)), comp(i2c, c)), comp(i2xm, c));
				const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
				if (A.w == 0.0f) {
					// cell outside every UPML box that shares a float4 chunk with the box; in place:
					// exactly one launch may update it -> the first box whose footprint holds it
					if (shell_earlier_box(p, sb.b, ic >> 2, j, k)) continue;
					setcomp(v0, c, leap(comp(v0, c), A.x, B.x, curl0));
					setcomp(v1, c, leap(comp(v1, c), A.y, B.y, curl1));
					setcomp(v2, c, leap(comp(v2, c), A.z, B.z, curl2));
				} else {
					if ((unsigned)li >= (unsigned)q.bn0) continue; // belongs to another box
					const float4 P0 = __ldg(p.tP0 + e[c]), P1 = __ldg(p.tP1 + e[c]), P2 = __ldg(p.tP2 + e[c]);
					const float* f = q.flux + frow + li;
					float* fo = q.flux_out + frow + li;
					setcomp(v0, c, leap_pml(comp(v0, c), A.x, B.x, curl0, P0.x, P1.x, P2.x, f, fo));
					setcomp(v1, c, leap_pml(comp(v1, c), A.y, B.y, curl1, P0.y, P1.y, P2.y, f + q.cs, fo + q.cs));
					setcomp(v2, c, leap_pml(comp(v2, c), A.z, B.z, curl2, P0.z, P1.z, P2.z, f + 2 * q.cs, fo + 2 * q.cs));
				}
				done |= 1u << c;
			}
			// cells the one-pass kernels skip: E_new goes to the destination set; in place (source set) only where a
			// one-pass kernel reads it as the +y / +z neighbour of its last row / plane
			const bool skipped = p.Xd2 && (k < p.sk0 || k >= p.sk1 || j < p.sjb || j >= p.sje);
			const bool inplace = !skipped || k == p.sk1 || j == p.sje;
			if (done == 15u) {
				if (inplace) {
					st4(p.Xd + o, v0);
					st4(p.Xd + p.comp + o, v1);
					st4(p.Xd + 2 * p.comp + o, v2);
				}
				if (skipped) {
					st4(p.Xd2 + o, v0);
					st4(p.Xd2 + p.comp + o, v1);
					st4(p.Xd2 + 2 * p.comp + o, v2);
				}
			} else {
#pragma unroll
				for (int c = 0; c < 4; ++c)
					if (done >> c & 1u) {
						if (inplace) {
							p.Xd[o + c] = comp(v0, c);
							p.Xd[p.comp + o + c] = comp(v1, c);
							p.Xd[2 * p.comp + o + c] = comp(v2, c);
						}
						if (skipped) {
							p.Xd2[o + c] = comp(v0, c);
							p.Xd2[p.comp + o + c] = comp(v1, c);
							p.Xd2[2 * p.comp + o + c] = comp(v2, c);
						}
					}
			}
		}
		i0km = i0c;
		i1km = i1c;
	}
}

template <typename IdxT>
__global__ void __launch_bounds__(256, SHELL_MIN_BLOCKS) k_shell_H(const __grid_constant__ ShellParams p)
{
	PDL_PROLOGUE();
	const ShellBlock sb = shell_block(p);
	const ShellBoxParams& q = p.box[sb.b];
	const int XL = q.xl;
	const int lane = threadIdx.x, sub = lane & (XL - 1);
	const int ch = sb.bx * XL + sub;
	const int lj = q.jr0 + (sb.by * blockDim.y + threadIdx.y) * (32 / XL) + lane / XL;
	const int kb = q.k0 + sb.bz * q.zchunk;
	const int ke = min(kb + q.zchunk, q.k1);
	if (kb >= ke) return;
	const bool active = ch < q.nchunk && lj < q.jr1;
	if (__all_sync(0xffffffffu, !active)) return;
	const int ic = (q.c0 + (ch < q.nchunk ? ch : 0)) * 4;
	const int j = q.bs1 + (lj < q.jr1 ? lj : q.jr0);
	const bool upd = active && j < p.ny - 1; // UpdateCurrents stops one line short (engine.cpp:179-183)
	const long long row = (long long)j * p.pitch + ic;
	const long long rowp = (long long)(j < p.ny - 1 ? j + 1 : j) * p.pitch + ic;
	const float* __restrict__ V0 = p.Y;
	const float* __restrict__ V1 = p.Y + p.comp;
	const float* __restrict__ V2 = p.Y + 2 * p.comp;
	const float* __restrict__ I0 = p.Xs;
	const float* __restrict__ I1 = p.Xs + p.comp;
	const float* __restrict__ I2 = p.Xs + 2 * p.comp;
	const bool has_right = ic + 4 < p.pitch;

	float4 v0c, v1c;
	{
		const long long o = (long long)kb * p.plane + row;
		v0c = ld4(V0 + o);
		v1c = ld4(V1 + o);
	}
	for (int k = kb; k < ke; ++k) {
		const long long o = (long long)k * p.plane + row;
		const long long op = (long long)k * p.plane + rowp;
		const long long on = o + p.plane; // plane k+1 is always held: k1 <= held planes - 1
		unsigned e[4];
		Idx4<IdxT>::load(p.idx, o, e);
		if (XL >= 16 && k + 1 < ke && active && (sub & 7) == 0) { // thin boxes: a line prefetch would fetch 128 B for the 48 B they use
			const long long of = on + p.plane;
			prefetch_l2(V0 + of); prefetch_l2(V1 + of); prefetch_l2(V2 + on);
			prefetch_l2(I0 + on); prefetch_l2(I1 + on); prefetch_l2(I2 + on);
			prefetch_l2(reinterpret_cast<const IdxT*>(p.idx) + on);
			const float* fn = q.flux + ((long long)(k + 1 - q.bs2) * q.bn1 + lj) * q.bn0 + max(ic - q.bs0, 0);
			prefetch_l2(fn); prefetch_l2(fn + q.cs); prefetch_l2(fn + 2 * q.cs);
		}
		const float4 v2c = ld4(V2 + o);
		const float4 v0n = ld4(V0 + on), v1n = ld4(V1 + on);
		const float4 v0jp = ld4(V0 + op), v2jp = ld4(V2 + op);
		float4 c0 = ld4(I0 + o), c1 = ld4(I1 + o), c2 = ld4(I2 + o);
		float r1 = __shfl_down_sync(0xffffffffu, v1c.x, 1, XL);
		float r2 = __shfl_down_sync(0xffffffffu, v2c.x, 1, XL);
		if (sub == XL - 1 || ch >= q.nchunk - 1) { // last lane of the group or last chunk of the box
			if (has_right) { r1 = V1[o + 4]; r2 = V2[o + 4]; }
			else { r1 = 0.0f; r2 = 0.0f; } // only reached by cells that are never written
		}
		const float4 v1xp = make_float4(v1c.y, v1c.z, v1c.w, r1);
		const float4 v2xp = make_float4(v2c.y, v2c.z, v2c.w, r2);
		if (upd) {
			const long long frow = ((long long)(k - q.bs2) * q.bn1 + lj) * q.bn0;
			unsigned done = 0;
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const int li = ic + c - q.bs0;
				if (ic + c >= p.nx - 1) continue;
				const float4 A = __ldg(p.tA + e[c]), B = __ldg(p.tB + e[c]);
				const float curl0 = fadd(fsub(fsub(comp(v2c, c), comp(v2jp, c)), comp(v1c, c)), comp(v1n, c));
				const float curl1 = fadd(fsub(fsub(comp(v0c, c), comp(v0n, c)), comp(v2c, c)), comp(v2xp, c));
				const float curl2 = fadd(fsub(fsub(comp(v1c, c), comp(v1xp, c)), comp(v0c, c)), comp(v0jp, c));
				if (A.w == 0.0f) {
					setcomp(c0, c, leap(comp(c0, c), A.x, B.x, curl0));
					setcomp(c1, c, leap(comp(c1, c), A.y, B.y, curl1));
					setcomp(c2, c, leap(comp(c2, c), A.z, B.z, curl2));
				} else {
					if ((unsigned)li >= (unsigned)q.bn0) continue; // belongs to another box
					const float4 P0 = __ldg(p.tP0 + e[c]), P1 = __ldg(p.tP1 + e[c]), P2 = __ldg(p.tP2 + e[c]);
					float* f = q.flux + frow + li;
					setcomp(c0, c, leap_pml(comp(c0, c), A.x, B.x, curl0, P0.x, P1.x, P2.x, f, f));
					setcomp(c1, c, leap_pml(comp(c1, c), A.y, B.y, curl1, P0.y, P1.y, P2.y, f + q.cs, f + q.cs));
					setcomp(c2, c, leap_pml(comp(c2, c), A.z, B.z, curl2, P0.z, P1.z, P2.z, f + 2 * q.cs, f + 2 * q.cs));
				}
				done |= 1u << c;
			}
			if (done == 15u) {
				st4(p.Xd + o, c0);
				st4(p.Xd + p.comp + o, c1);
				st4(p.Xd + 2 * p.comp + o, c2);
			} else {
#pragma unroll
				for (int c = 0; c < 4; ++c)
					if (done >> c & 1u) {
						p.Xd[o + c] = comp(c0, c);
						p.Xd[p.comp + o + c] = comp(c1, c);
						p.Xd[2 * p.comp + o + c] = comp(c2, c);
					}
			}
		}
		v0c = v0n;
		v1c = v1n;
	}
}

// plain device-to-device plane copy helper for the parts of the destination set the fused
// kernel does not write (ghost planes are filled by the halo pushes)
__global__ void k_copy_f4(const float4* __restrict__ src, float4* __restrict__ dst, long long n4)
{
	PDL_PROLOGUE();
	for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) dst[q] = src[q];
}
