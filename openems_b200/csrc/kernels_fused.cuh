// kernels_fused.cuh -- one-pass timestep: E half-step and H half-step fused into one kernel.
//
// Traffic per cell and timestep drops from 76 B (two passes: each reads E and H and writes one
// of them) to 50 B (read E, H and the operator index once, write E and H once).  The arithmetic
// per cell is exactly that of k_update_E / k_update_H (same helpers, same roundings), only the
// order in which cells are visited changes, so the results are bit-identical.
//
// Scheme (z-marching, out of place: fields and UPML flux are ping-pong buffers, the source set
// is never written during a step):
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
#pragma once
#include "kernels.cuh"

#ifndef FUSED_TY
#define FUSED_TY 7   // rows per tile; +1 halo-row warp = 8 warps = 256 threads, so 2 blocks/SM may use 128 registers
#endif

struct FusedParams {
	const float* Vs; const float* Is;   // source set (timestep n)
	float* Vd; float* Id;               // destination set (timestep n+1)
	const float* fVs; float* fVd;       // UPML voltage flux, source / destination
	const float* fIs; float* fId;       // UPML current flux
	const void* idx;
	const float4 *eA, *eB, *eP0, *eP1, *eP2; // E tables (vv|flag, vi, aux vv, vvfn, vvfo)
	const float4 *hA, *hB, *hP0, *hP1, *hP2; // H tables
	int nx, ny, nz;      // nz = local planes held
	int pitch;
	long long plane, comp;
	int kE0, kE1;        // local planes whose E this kernel produces (owned planes)
	int kH0, kH1;        // local planes whose H is UPDATED here; other owned planes are copied through
	int kHc1;            // H planes [kH1, kHc1) are copied through (top of the domain); [kHc1, kE1) left to the slab kernel
	int zchunk;
	int nboxes;
	PmlBox box[OEMS_MAX_PML_BOXES];
};

template <typename P>
__device__ __forceinline__ long long pml_offset_any(const P& p, int i, int j, int k, long long& cs)
{
#pragma unroll 1
	for (int b = 0; b < p.nboxes; ++b) {
		const PmlBox& B = p.box[b];
		const int li = i - B.s[0], lj = j - B.s[1], lk = k - B.s[2];
		if ((unsigned)li < (unsigned)B.n[0] && (unsigned)lj < (unsigned)B.n[1] && (unsigned)lk < (unsigned)B.n[2]) {
			cs = (long long)B.n[0] * B.n[1] * B.n[2];
			return B.off + ((long long)lk * B.n[1] + lj) * B.n[0] + li;
		}
	}
	return -1;
}

// out-of-place UPML update of one component: reads the flux of timestep n, returns the new
// field value and the new flux (engine_ext_upml.cpp:52-137 / :144-229 around the leapfrog)
__device__ __forceinline__ float leap_pml_oop(float X, float m_vv, float m_vi, float curl, float a_vv, float a_fn,
                                              float a_fo, float F, float& Fn)
{
	const float f = fsub(fmul(a_vv, X), fmul(a_fo, F));
	Fn = fadd(fmul(F, m_vv), fmul(m_vi, curl));
	return fadd(f, fmul(a_fn, Fn));
}

// all three components of one cell; flux_out == nullptr: do not store the new flux (halo cells)
template <bool HAS_PML, typename P>
__device__ __forceinline__ void cell_update(const P& p, const float4* tP0, const float4* tP1, const float4* tP2,
                                            const float* flux_in, float* flux_out, unsigned e, const float4& A,
                                            const float4& B, int x, int j, int k, float curl0, float curl1, float curl2,
                                            float& x0, float& x1, float& x2)
{
	if (HAS_PML && A.w != 0.0f) {
		long long cs;
		const long long fo = pml_offset_any(p, x, j, k, cs);
		if (fo >= 0) {
			const float4 P0 = __ldg(tP0 + e), P1 = __ldg(tP1 + e), P2 = __ldg(tP2 + e);
			float f0, f1, f2;
			x0 = leap_pml_oop(x0, A.x, B.x, curl0, P0.x, P1.x, P2.x, flux_in[fo], f0);
			x1 = leap_pml_oop(x1, A.y, B.y, curl1, P0.y, P1.y, P2.y, flux_in[fo + cs], f1);
			x2 = leap_pml_oop(x2, A.z, B.z, curl2, P0.z, P1.z, P2.z, flux_in[fo + 2 * cs], f2);
			if (flux_out) { flux_out[fo] = f0; flux_out[fo + cs] = f1; flux_out[fo + 2 * cs] = f2; }
		}
		return;
	}
	x0 = leap(x0, A.x, B.x, curl0);
	x1 = leap(x1, A.y, B.y, curl1);
	x2 = leap(x2, A.z, B.z, curl2);
}

// flux offset like pml_offset_any, also returning the box's plane stride (for the L2 prefetch of
// the next plane) -- used by the batched UPML pass of k_fused_EH
template <typename P>
__device__ __forceinline__ long long pml_offset_ps(const P& p, int i, int j, int k, long long& cs, long long& ps)
{
#pragma unroll 1
	for (int b = 0; b < p.nboxes; ++b) {
		const PmlBox& B = p.box[b];
		const int li = i - B.s[0], lj = j - B.s[1], lk = k - B.s[2];
		if ((unsigned)li < (unsigned)B.n[0] && (unsigned)lj < (unsigned)B.n[1] && (unsigned)lk < (unsigned)B.n[2]) {
			ps = (long long)B.n[0] * B.n[1];
			cs = ps * B.n[2];
			if (lk + 1 >= B.n[2]) ps = 0;
			return B.off + ((long long)lk * B.n[1] + lj) * B.n[0] + li;
		}
	}
	return -1;
}

template <typename IdxT, bool HAS_PML>
__global__ void __launch_bounds__(32 * (FUSED_TY + 1), 2) k_fused_EH(const __grid_constant__ FusedParams p)
{
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
		float4 v0 = ld4(V0 + o), v1 = ld4(V1 + o), v2 = ld4(V2 + o);
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
			const bool st = !halo_row && kk < ke;
			unsigned pm = 0; // UPML cells of this thread
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				if (c > 0 && !uni) { A = __ldg(p.eA + e[c]); B = __ldg(p.eB + e[c]); }
				if (HAS_PML && A.w != 0.0f) { pm |= 1u << c; continue; }
				const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(hk1, c));
				const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(hk0, c)), comp(i2c, c)), comp(i2xm, c));
				const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
				setcomp(v0, c, leap(comp(v0, c), A.x, B.x, curl0));
				setcomp(v1, c, leap(comp(v1, c), A.y, B.y, curl1));
				setcomp(v2, c, leap(comp(v2, c), A.z, B.z, curl2));
			}
			if (HAS_PML && pm) {
				// rare path: all flux loads of the thread's UPML cells are issued before the first use
				long long fo[4], cs[4];
				float F[4][3];
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					fo[c] = -1;
					if (pm >> c & 1u) {
						long long ps;
						fo[c] = pml_offset_ps(p, ic + c, jc, kk, cs[c], ps);
						if (fo[c] >= 0) {
							F[c][0] = __ldg(p.fVs + fo[c]); F[c][1] = __ldg(p.fVs + fo[c] + cs[c]); F[c][2] = __ldg(p.fVs + fo[c] + 2 * cs[c]);
							if (ps && (c == 0 || c == 3)) { prefetch_l2(p.fVs + fo[c] + ps); prefetch_l2(p.fVs + fo[c] + cs[c] + ps); prefetch_l2(p.fVs + fo[c] + 2 * cs[c] + ps); }
						}
					}
				}
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					if (fo[c] < 0) continue;
					const float4 Ac = __ldg(p.eA + e[c]), Bc = __ldg(p.eB + e[c]);
					const float4 P0 = __ldg(p.eP0 + e[c]), P1 = __ldg(p.eP1 + e[c]), P2 = __ldg(p.eP2 + e[c]);
					const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(hk1, c));
					const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(hk0, c)), comp(i2c, c)), comp(i2xm, c));
					const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
					float f;
					setcomp(v0, c, leap_pml_oop(comp(v0, c), Ac.x, Bc.x, curl0, P0.x, P1.x, P2.x, F[c][0], f)); F[c][0] = f;
					setcomp(v1, c, leap_pml_oop(comp(v1, c), Ac.y, Bc.y, curl1, P0.y, P1.y, P2.y, F[c][1], f)); F[c][1] = f;
					setcomp(v2, c, leap_pml_oop(comp(v2, c), Ac.z, Bc.z, curl2, P0.z, P1.z, P2.z, F[c][2], f)); F[c][2] = f;
				}
				if (st) {
#pragma unroll
					for (int c = 0; c < 4; ++c)
						if (fo[c] >= 0) { p.fVd[fo[c]] = F[c][0]; p.fVd[fo[c] + cs[c]] = F[c][1]; p.fVd[fo[c] + 2 * cs[c]] = F[c][2]; }
				}
			}
			if (st) {
				st4(p.Vd + o, v0);
				st4(p.Vd + p.comp + o, v1);
				st4(p.Vd + 2 * p.comp + o, v2);
			}
			if (hcol) {
				// V1, V2 of cell (xe, j, kk): engine.cpp:148-166 with the x-1 neighbour = my last cell
				const unsigned ex = reinterpret_cast<const IdxT*>(p.idx)[o + 4];
				const float4 Ax = __ldg(p.eA + ex), Bx = __ldg(p.eB + ex);
				const float xi0 = I0[o + 4], xi1 = I1[o + 4], xi2 = I2[o + 4], xi0jm = I0[om + 4], xi2jm = I2[om + 4];
				const float c0x = fadd(fsub(fsub(xi2, xi2jm), xi1), 0.0f); // V0 of the halo cell is not needed
				const float c1x = fadd(fsub(fsub(xi0, hcI0), xi2), i2c.w);
				const float c2x = fadd(fsub(fsub(xi1, i1c.w), xi0), xi0jm);
				float a = 0.0f, b = V1[o + 4], d = V2[o + 4];
				cell_update<HAS_PML>(p, p.eP0, p.eP1, p.eP2, p.fVs, (float*)nullptr, ex, Ax, Bx, xe, jc, kk, c0x, c1x, c2x, a, b, d);
				nV1 = b; nV2 = d; nI0 = xi0;
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
			if (k < he && j < p.ny - 1) {
				const float4 v0jp = xV0[k % 3][ty + 1][lane], v2jp = xV2[k % 3][ty + 1][lane];
				const float4 v1xp = make_float4(ek1.y, ek1.z, ek1.w, r1);
				const float4 v2xp = make_float4(ek2.y, ek2.z, ek2.w, r2);
				const bool uni = (ek_idx[0] == ek_idx[1]) & (ek_idx[1] == ek_idx[2]) & (ek_idx[2] == ek_idx[3]);
				float4 A = __ldg(p.hA + ek_idx[0]), B = __ldg(p.hB + ek_idx[0]);
				unsigned pm = 0;
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					if (c > 0 && !uni) { A = __ldg(p.hA + ek_idx[c]); B = __ldg(p.hB + ek_idx[c]); }
					if (ic + c < p.nx - 1) {
						if (HAS_PML && A.w != 0.0f) { pm |= 1u << c; continue; }
						const float curl0 = fadd(fsub(fsub(comp(ek2, c), comp(v2jp, c)), comp(ek1, c)), comp(v1, c));
						const float curl1 = fadd(fsub(fsub(comp(ek0, c), comp(v0, c)), comp(ek2, c)), comp(v2xp, c));
						const float curl2 = fadd(fsub(fsub(comp(ek1, c), comp(v1xp, c)), comp(ek0, c)), comp(v0jp, c));
						setcomp(c0, c, leap(comp(c0, c), A.x, B.x, curl0));
						setcomp(c1, c, leap(comp(c1, c), A.y, B.y, curl1));
						setcomp(c2, c, leap(comp(c2, c), A.z, B.z, curl2));
					}
				}
				if (HAS_PML && pm) {
					long long fo[4], cs[4];
					float F[4][3];
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						fo[c] = -1;
						if (pm >> c & 1u) {
							long long ps;
							fo[c] = pml_offset_ps(p, ic + c, jc, k, cs[c], ps);
							if (fo[c] >= 0) {
								F[c][0] = __ldg(p.fIs + fo[c]); F[c][1] = __ldg(p.fIs + fo[c] + cs[c]); F[c][2] = __ldg(p.fIs + fo[c] + 2 * cs[c]);
								if (ps && (c == 0 || c == 3)) { prefetch_l2(p.fIs + fo[c] + ps); prefetch_l2(p.fIs + fo[c] + cs[c] + ps); prefetch_l2(p.fIs + fo[c] + 2 * cs[c] + ps); }
							}
						}
					}
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						if (fo[c] < 0) continue;
						const float4 Ac = __ldg(p.hA + ek_idx[c]), Bc = __ldg(p.hB + ek_idx[c]);
						const float4 P0 = __ldg(p.hP0 + ek_idx[c]), P1 = __ldg(p.hP1 + ek_idx[c]), P2 = __ldg(p.hP2 + ek_idx[c]);
						const float curl0 = fadd(fsub(fsub(comp(ek2, c), comp(v2jp, c)), comp(ek1, c)), comp(v1, c));
						const float curl1 = fadd(fsub(fsub(comp(ek0, c), comp(v0, c)), comp(ek2, c)), comp(v2xp, c));
						const float curl2 = fadd(fsub(fsub(comp(ek1, c), comp(v1xp, c)), comp(ek0, c)), comp(v0jp, c));
						float f;
						setcomp(c0, c, leap_pml_oop(comp(c0, c), Ac.x, Bc.x, curl0, P0.x, P1.x, P2.x, F[c][0], f)); F[c][0] = f;
						setcomp(c1, c, leap_pml_oop(comp(c1, c), Ac.y, Bc.y, curl1, P0.y, P1.y, P2.y, F[c][1], f)); F[c][1] = f;
						setcomp(c2, c, leap_pml_oop(comp(c2, c), Ac.z, Bc.z, curl2, P0.z, P1.z, P2.z, F[c][2], f)); F[c][2] = f;
					}
#pragma unroll
					for (int c = 0; c < 4; ++c)
						if (fo[c] >= 0) { p.fId[fo[c]] = F[c][0]; p.fId[fo[c] + cs[c]] = F[c][1]; p.fId[fo[c] + 2 * cs[c]] = F[c][2]; }
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
// k_fused_EH ran.  One thread per listed cell, all three components.
// ---------------------------------------------------------------------------------------
struct FixParams {
	const float* Is; float* Id;
	const float* Vd;                 // final E of this timestep
	const float* fIs; float* fId;
	const void* idx;
	const float4 *hA, *hB, *hP0, *hP1, *hP2;
	const int* cell;                 // [count][3] x, y, local z
	long long count;
	int nx, ny;
	int pitch; long long plane, comp;
	int nboxes;
	PmlBox box[OEMS_MAX_PML_BOXES];
};

template <typename IdxT, bool HAS_PML>
__global__ void k_fix_H(const __grid_constant__ FixParams p)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	const int x = p.cell[3 * t], j = p.cell[3 * t + 1], k = p.cell[3 * t + 2];
	const long long o = (long long)k * p.plane + (long long)j * p.pitch + x;
	const float* V0 = p.Vd; const float* V1 = p.Vd + p.comp; const float* V2 = p.Vd + 2 * p.comp;
	const unsigned e = reinterpret_cast<const IdxT*>(p.idx)[o];
	const float4 A = __ldg(p.hA + e), B = __ldg(p.hB + e);
	const float v0 = V0[o], v1 = V1[o], v2 = V2[o];
	const float curl0 = fadd(fsub(fsub(v2, V2[o + p.pitch]), v1), V1[o + p.plane]);
	const float curl1 = fadd(fsub(fsub(v0, V0[o + p.plane]), v2), V2[o + 1]);
	const float curl2 = fadd(fsub(fsub(v1, V1[o + 1]), v0), V0[o + p.pitch]);
	float a = p.Is[o], b = p.Is[p.comp + o], d = p.Is[2 * p.comp + o];
	cell_update<HAS_PML>(p, p.hP0, p.hP1, p.hP2, p.fIs, p.fId, e, A, B, x, j, k, curl0, curl1, curl2, a, b, d);
	p.Id[o] = a; p.Id[p.comp + o] = b; p.Id[2 * p.comp + o] = d;
}

// plain device-to-device plane copy helper for the parts of the destination set the fused
// kernel does not write (ghost planes are filled by the halo pushes)
__global__ void k_copy_f4(const float4* __restrict__ src, float4* __restrict__ dst, long long n4)
{
	for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) dst[q] = src[q];
}
