// kernels.cuh -- sm_100a device kernels of the openEMS FDTD hot path.
//
// Arithmetic contract (parity with the reference's sse-compressed engine, bit for bit):
// fp32, every multiply and add rounded separately (explicit __fmul_rn/__fadd_rn, and the
// library is compiled -fmad=false), left-to-right sums exactly as written in
// FDTD/engine.cpp:110-222, denormals flushed (-ftz=true == tools/denormal.h:19-30).
//
// Data layout in HBM: fields V[3][nz][ny][pitch], I[3][nz][ny][pitch], x contiguous,
// pitch = nx rounded up to 32 floats (rows start on 128 B lines); one operator index per
// cell idx[nz][ny][pitch] (u16 or u32) into de-duplicated float4 coefficient tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Programmatic dependent launch (Engine::launch_k): a kernel may be set up while its predecessor in the stream is still
// running.  First statement of EVERY kernel: let the successor's launch begin (small grids only: blocks of a successor
// that sit on an SM waiting would take slots from a predecessor that still has blocks to schedule), then wait until the
// predecessor grid has completed and its memory operations are visible.  Without the launch attribute both are no-ops.
#define PDL_EARLY_MAX_BLOCKS 1184   // 8 x 148
#define PDL_PROLOGUE()                                                                                         \
	do {                                                                                                       \
		if (gridDim.x * gridDim.y * gridDim.z <= PDL_EARLY_MAX_BLOCKS) asm volatile("griddepcontrol.launch_dependents;"); \
		asm volatile("griddepcontrol.wait;" ::: "memory");                                                     \
	} while (0)

#define OEMS_MAX_PML_BOXES 8
// tuning macros (overridable from the nvcc command line for sweeps)
#ifndef OEMS_MIN_BLOCKS
#define OEMS_MIN_BLOCKS 4
#endif
#ifndef OEMS_PREFETCH_DIST
#define OEMS_PREFETCH_DIST 1   // planes ahead that are pulled into L2 (0 = off); 1 measured best (profiles/)
#endif
#ifndef OEMS_PREFETCH_LANES
#define OEMS_PREFETCH_LANES 7  // lane mask: (lane & mask)==0 issues the prefetch of its 128 B line
#endif
#define OEMS_MAX_MUR 6

struct PmlBox {
	int s[3];       // start (x, y, local z)
	int n[3];       // lines
	long long off;  // offset of this box inside the flux buffer (floats); box layout [c][k][j][i]
};

struct StencilParams {
	float* V;
	float* I;
	const void* idx;
	// E step: tA = {vv0,vv1,vv2,pml flag}, tB = {vi0,vi1,vi2,0}, tP0..2 = aux vv, vvfn, vvfo
	// H step: tA = {ii0,ii1,ii2,pml flag}, tB = {iv0,iv1,iv2,0}, tP0..2 = aux ii, iifn, iifo
	const float4* tA;
	const float4* tB;
	const float4* tP0;
	const float4* tP1;
	const float4* tP2;
	float* flux;
	// out-of-place mode of k_update_H (slab top plane of the fused schedule): results go to Iout /
	// flux_out, every cell of the planes is written (cells outside the update range are copied)
	float* Iout;
	float* flux_out;
	int nx, ny, nz;      // lines held by this GPU (nz = local planes incl. ghosts)
	int pitch;
	long long plane;     // pitch*ny
	long long comp;      // plane*nz
	int k0, k1;          // local plane range to update
	int zchunk;
	unsigned* tick;      // k_small_H: the timestep counter, advanced by this launch when it is the last kernel of the timestep (else NULL)
	int nboxes;
	PmlBox box[OEMS_MAX_PML_BOXES];
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// streaming variants for data touched once per half-step (the field being updated)
#ifdef OEMS_STREAM_RW
__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4s(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }
#else
__device__ __forceinline__ float4 ld4s(const float* p) { return ld4(p); }
__device__ __forceinline__ void st4s(float* p, const float4& v) { st4(p, v); }
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename IdxT> struct Idx4;
template <> struct Idx4<uint16_t> {
	__device__ __forceinline__ static void load(const void* base, long long off, unsigned e[4])
	{
		ushort4 v = *reinterpret_cast<const ushort4*>(reinterpret_cast<const uint16_t*>(base) + off);
		e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
	}
};
template <> struct Idx4<uint32_t> {
	__device__ __forceinline__ static void load(const void* base, long long off, unsigned e[4])
	{
		uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(base) + off);
		e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
	}
};

__device__ __forceinline__ float comp(const float4& v, int c)
{
	return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}
__device__ __forceinline__ void setcomp(float4& v, int c, float x)
{
	if (c == 0) v.x = x; else if (c == 1) v.y = x; else if (c == 2) v.z = x; else v.w = x;
}

// flux offset of cell (i,j,k) if it lies in a UPML box, -1 otherwise; cs = component stride
__device__ __forceinline__ long long pml_flux_offset(const StencilParams& p, int i, int j, int k, long long& cs)
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

// One field component of one cell: plain leapfrog (engine.cpp:137-146) or, inside a UPML box,
// the fused sequence pre-update / leapfrog / post-update of Engine_Ext_UPML
// (engine_ext_upml.cpp:52-137 for voltages, :144-229 for currents):
//   f  = a_vv*X - a_fo*F ;  F' = F*m_vv + m_vi*curl ;  X' = f + a_fn*F'
__device__ __forceinline__ float leap(float X, float m_vv, float m_vi, float curl)
{
	return fadd(fmul(X, m_vv), fmul(m_vi, curl));
}
__device__ __forceinline__ float leap_pml(float X, float m_vv, float m_vi, float curl, float a_vv,
                                          float a_fn, float a_fo, const float* flux_in, float* flux_out)
{
	const float F = *flux_in;
	const float f = fsub(fmul(a_vv, X), fmul(a_fo, F));
	const float Fn = fadd(fmul(F, m_vv), fmul(m_vi, curl));
	*flux_out = Fn;
	return fadd(f, fmul(a_fn, Fn));
}

// ---------------------------------------------------------------------------------------
// E half-step: Engine::UpdateVoltages engine.cpp:110-168 (+ fused UPML hooks).
// block = (32, rows): a warp owns 128 consecutive x cells of one row (float4 per lane) and
// marches zchunk planes in z, carrying the k-1 plane of I0/I1 in registers; the j-1 row comes
// through L1/L2 (it is the row the warp above just loaded), the i-1 element by warp shuffle.
// ---------------------------------------------------------------------------------------
template <typename IdxT, bool HAS_PML>
__global__ void __launch_bounds__(256, OEMS_MIN_BLOCKS) k_update_E(const __grid_constant__ StencilParams p)
{
	PDL_PROLOGUE();
	const int lane = threadIdx.x;
	const int i0 = (blockIdx.x * 32 + lane) * 4;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int kb = p.k0 + blockIdx.z * p.zchunk;
	const int ke = min(kb + p.zchunk, p.k1);
	if (j >= p.ny || kb >= ke) return;
	const bool active = i0 < p.pitch;
	const int ic = active ? i0 : 0;
	const int jm = j - (j > 0);
	const long long row = (long long)j * p.pitch + ic;
	const long long rowm = (long long)jm * p.pitch + ic;
	const float* __restrict__ I0 = p.I;
	const float* __restrict__ I1 = p.I + p.comp;
	const float* __restrict__ I2 = p.I + 2 * p.comp;
	float* V0 = p.V;
	float* V1 = p.V + p.comp;
	float* V2 = p.V + 2 * p.comp;

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
#if OEMS_PREFETCH_DIST > 0
		if (k + OEMS_PREFETCH_DIST < ke && (lane & OEMS_PREFETCH_LANES) == 0) {
			const long long of = o + (long long)OEMS_PREFETCH_DIST * p.plane;
			prefetch_l2(I0 + of); prefetch_l2(I1 + of); prefetch_l2(I2 + of);
			prefetch_l2(V0 + of); prefetch_l2(V1 + of); prefetch_l2(V2 + of);
			if ((lane & 15) == 0) prefetch_l2(reinterpret_cast<const IdxT*>(p.idx) + of);
		}
#endif
		const float4 i0c = ld4(I0 + o), i1c = ld4(I1 + o), i2c = ld4(I2 + o);
		const float4 i0jm = ld4(I0 + om), i2jm = ld4(I2 + om);
		float4 v0 = ld4s(V0 + o), v1 = ld4s(V1 + o), v2 = ld4s(V2 + o);
		// i-1 neighbours of I1, I2
		float l1 = __shfl_up_sync(0xffffffffu, i1c.w, 1);
		float l2 = __shfl_up_sync(0xffffffffu, i2c.w, 1);
		if (lane == 0) {
			if (ic > 0) { l1 = I1[o - 1]; l2 = I2[o - 1]; }
			else { l1 = i1c.x; l2 = i2c.x; }
		}
		const float4 i1xm = make_float4(l1, i1c.x, i1c.y, i1c.z);
		const float4 i2xm = make_float4(l2, i2c.x, i2c.y, i2c.z);
		if (active) {
			const bool uni = (e[0] == e[1]) & (e[1] == e[2]) & (e[2] == e[3]);
			float4 A = __ldg(p.tA + e[0]), B = __ldg(p.tB + e[0]);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				if (c > 0 && !uni) { A = __ldg(p.tA + e[c]); B = __ldg(p.tB + e[c]); }
				// ((a - b) - c) + d, engine.cpp:139-144
				const float curl0 = fadd(fsub(fsub(comp(i2c, c), comp(i2jm, c)), comp(i1c, c)), comp(i1km, c));
				const float curl1 = fadd(fsub(fsub(comp(i0c, c), comp(i0km, c)), comp(i2c, c)), comp(i2xm, c));
				const float curl2 = fadd(fsub(fsub(comp(i1c, c), comp(i1xm, c)), comp(i0c, c)), comp(i0jm, c));
				if (HAS_PML && A.w != 0.0f) {
					long long cs;
					const long long fo = pml_flux_offset(p, ic + c, j, k, cs);
					const float4 P0 = __ldg(p.tP0 + e[c]), P1 = __ldg(p.tP1 + e[c]), P2 = __ldg(p.tP2 + e[c]);
					if (fo >= 0) {
						setcomp(v0, c, leap_pml(comp(v0, c), A.x, B.x, curl0, P0.x, P1.x, P2.x, p.flux + fo, p.flux + fo));
						setcomp(v1, c, leap_pml(comp(v1, c), A.y, B.y, curl1, P0.y, P1.y, P2.y, p.flux + fo + cs, p.flux + fo + cs));
						setcomp(v2, c, leap_pml(comp(v2, c), A.z, B.z, curl2, P0.z, P1.z, P2.z, p.flux + fo + 2 * cs, p.flux + fo + 2 * cs));
					}
				} else {
					setcomp(v0, c, leap(comp(v0, c), A.x, B.x, curl0));
					setcomp(v1, c, leap(comp(v1, c), A.y, B.y, curl1));
					setcomp(v2, c, leap(comp(v2, c), A.z, B.z, curl2));
				}
			}
			st4s(V0 + o, v0);
			st4s(V1 + o, v1);
			st4s(V2 + o, v2);
		}
		i0km = i0c;
		i1km = i1c;
	}
}

// ---------------------------------------------------------------------------------------
// H half-step: Engine::UpdateCurrents engine.cpp:170-222 (+ fused UPML hooks); cells with
// i < nx-1, j < ny-1, global k < nz-1 only.  Marches z upward carrying plane k+1 of V0/V1.
// ---------------------------------------------------------------------------------------
template <typename IdxT, bool HAS_PML>
__global__ void __launch_bounds__(256, OEMS_MIN_BLOCKS) k_update_H(const __grid_constant__ StencilParams p)
{
	PDL_PROLOGUE();
	const int lane = threadIdx.x;
	const int i0 = (blockIdx.x * 32 + lane) * 4;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int kb = p.k0 + blockIdx.z * p.zchunk;
	const int ke = min(kb + p.zchunk, p.k1);
	const bool oop = p.Iout != nullptr;
	if (j >= p.ny - (oop ? 0 : 1) || kb >= ke) return;
	const bool upd = j < p.ny - 1; // the last row is only copied (out-of-place mode)
	const bool active = i0 < p.pitch;
	const int ic = active ? i0 : 0;
	const long long row = (long long)j * p.pitch + ic;
	const long long rowp = (long long)(upd ? j + 1 : j) * p.pitch + ic;
	const float* __restrict__ V0 = p.V;
	const float* __restrict__ V1 = p.V + p.comp;
	const float* __restrict__ V2 = p.V + 2 * p.comp;
	const float* I0 = p.I;
	const float* I1 = p.I + p.comp;
	const float* I2 = p.I + 2 * p.comp;
	float* O0 = oop ? p.Iout : p.I;
	float* O1 = O0 + p.comp;
	float* O2 = O0 + 2 * p.comp;
	float* fout = (oop && p.flux_out) ? p.flux_out : p.flux;
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
		const long long on = o + p.plane;
		unsigned e[4];
		Idx4<IdxT>::load(p.idx, o, e);
#if OEMS_PREFETCH_DIST > 0
		if (k + OEMS_PREFETCH_DIST < ke && (lane & OEMS_PREFETCH_LANES) == 0) {
			const long long of = on + (long long)OEMS_PREFETCH_DIST * p.plane;
			prefetch_l2(V0 + of); prefetch_l2(V1 + of); prefetch_l2(V2 + of - p.plane);
			prefetch_l2(I0 + of - p.plane); prefetch_l2(I1 + of - p.plane); prefetch_l2(I2 + of - p.plane);
			if ((lane & 15) == 0) prefetch_l2(reinterpret_cast<const IdxT*>(p.idx) + of - p.plane);
		}
#endif
		const float4 v2c = ld4(V2 + o);
		const float4 v0n = ld4(V0 + on), v1n = ld4(V1 + on);
		const float4 v0jp = ld4(V0 + op), v2jp = ld4(V2 + op);
		float4 c0 = ld4s(I0 + o), c1 = ld4s(I1 + o), c2 = ld4s(I2 + o);
		float r1 = __shfl_down_sync(0xffffffffu, v1c.x, 1);
		float r2 = __shfl_down_sync(0xffffffffu, v2c.x, 1);
		if (lane == 31) {
			if (has_right) { r1 = V1[o + 4]; r2 = V2[o + 4]; }
			else { r1 = 0.0f; r2 = 0.0f; } // only reached by cells that are never written
		}
		const float4 v1xp = make_float4(v1c.y, v1c.z, v1c.w, r1);
		const float4 v2xp = make_float4(v2c.y, v2c.z, v2c.w, r2);
		if (active) {
			const bool uni = (e[0] == e[1]) & (e[1] == e[2]) & (e[2] == e[3]);
			float4 A = __ldg(p.tA + e[0]), B = __ldg(p.tB + e[0]);
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				if (c > 0 && !uni) { A = __ldg(p.tA + e[c]); B = __ldg(p.tB + e[c]); }
				if (upd && ic + c < p.nx - 1) {
					const float curl0 = fadd(fsub(fsub(comp(v2c, c), comp(v2jp, c)), comp(v1c, c)), comp(v1n, c));
					const float curl1 = fadd(fsub(fsub(comp(v0c, c), comp(v0n, c)), comp(v2c, c)), comp(v2xp, c));
					const float curl2 = fadd(fsub(fsub(comp(v1c, c), comp(v1xp, c)), comp(v0c, c)), comp(v0jp, c));
					if (HAS_PML && A.w != 0.0f) {
						long long cs;
						const long long fo = pml_flux_offset(p, ic + c, j, k, cs);
						const float4 P0 = __ldg(p.tP0 + e[c]), P1 = __ldg(p.tP1 + e[c]), P2 = __ldg(p.tP2 + e[c]);
						if (fo >= 0) {
							setcomp(c0, c, leap_pml(comp(c0, c), A.x, B.x, curl0, P0.x, P1.x, P2.x, p.flux + fo, fout + fo));
							setcomp(c1, c, leap_pml(comp(c1, c), A.y, B.y, curl1, P0.y, P1.y, P2.y, p.flux + fo + cs, fout + fo + cs));
							setcomp(c2, c, leap_pml(comp(c2, c), A.z, B.z, curl2, P0.z, P1.z, P2.z, p.flux + fo + 2 * cs, fout + fo + 2 * cs));
						}
					} else {
						setcomp(c0, c, leap(comp(c0, c), A.x, B.x, curl0));
						setcomp(c1, c, leap(comp(c1, c), A.y, B.y, curl1));
						setcomp(c2, c, leap(comp(c2, c), A.z, B.z, curl2));
					}
				}
			}
			st4s(O0 + o, c0);
			st4s(O1 + o, c1);
			st4s(O2 + o, c2);
		}
		v0c = v0n;
		v1c = v1n;
	}
}

// ---------------------------------------------------------------------------------------
// Small meshes (a few 10^5 cells: the tutorial configs C1-C3): the same two half-steps with ONE CELL PER THREAD.
// k_update_E / k_update_H give every thread four cells and a z march -- the right shape to stream a large mesh,
// but a 196 k cell mesh then runs on 10 warps per SM and a timestep half costs 8-13 us of exposed load latency
// (profiles/experiments_r02.md #10).  Here a thread updates one cell of one plane: four times the warps, all loads
// of a cell independent of each other except the coefficient gather behind the index load; the neighbours come
// through L1/L2 (the mesh fits L2 many times over).  Same helpers, same operand order: bit-identical.
// block = (32, rows), grid = (ceil(nx / 32), ceil(ny / rows), planes).
// ---------------------------------------------------------------------------------------
template <typename IdxT, bool HAS_PML>
__global__ void __launch_bounds__(256) k_small_E(const __grid_constant__ StencilParams p)
{
	PDL_PROLOGUE();
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = p.k0 + blockIdx.z;
	if (i >= p.nx || j >= p.ny || k >= p.k1) return;
	const long long o = (long long)k * p.plane + (long long)j * p.pitch + i;
	const long long om = o - (j > 0 ? p.pitch : 0);
	const long long ok = o - (k > 0 ? p.plane : 0);   // z-1 clamp only at the bottom of the (local) domain
	const long long ox = o - (i > 0 ? 1 : 0);
	const float* __restrict__ I0 = p.I;
	const float* __restrict__ I1 = p.I + p.comp;
	const float* __restrict__ I2 = p.I + 2 * p.comp;
	const unsigned e = reinterpret_cast<const IdxT*>(p.idx)[o];
	const float i0c = I0[o], i1c = I1[o], i2c = I2[o];
	const float i0jm = I0[om], i2jm = I2[om];
	const float i0km = I0[ok], i1km = I1[ok];
	const float i1xm = I1[ox], i2xm = I2[ox];
	float v0 = p.V[o], v1 = p.V[p.comp + o], v2 = p.V[2 * p.comp + o];
	const float4 A = __ldg(p.tA + e), B = __ldg(p.tB + e);
	// ((a - b) - c) + d, engine.cpp:139-144
	const float curl0 = fadd(fsub(fsub(i2c, i2jm), i1c), i1km);
	const float curl1 = fadd(fsub(fsub(i0c, i0km), i2c), i2xm);
	const float curl2 = fadd(fsub(fsub(i1c, i1xm), i0c), i0jm);
	if (HAS_PML && A.w != 0.0f) {
		long long cs;
		const long long fo = pml_flux_offset(p, i, j, k, cs);
		if (fo < 0) return; // flagged cell outside the boxes held here: left alone, as in k_update_E
		const float4 P0 = __ldg(p.tP0 + e), P1 = __ldg(p.tP1 + e), P2 = __ldg(p.tP2 + e);
		v0 = leap_pml(v0, A.x, B.x, curl0, P0.x, P1.x, P2.x, p.flux + fo, p.flux + fo);
		v1 = leap_pml(v1, A.y, B.y, curl1, P0.y, P1.y, P2.y, p.flux + fo + cs, p.flux + fo + cs);
		v2 = leap_pml(v2, A.z, B.z, curl2, P0.z, P1.z, P2.z, p.flux + fo + 2 * cs, p.flux + fo + 2 * cs);
	} else {
		v0 = leap(v0, A.x, B.x, curl0);
		v1 = leap(v1, A.y, B.y, curl1);
		v2 = leap(v2, A.z, B.z, curl2);
	}
	p.V[o] = v0; p.V[p.comp + o] = v1; p.V[2 * p.comp + o] = v2;
}

template <typename IdxT, bool HAS_PML>
__global__ void __launch_bounds__(256) k_small_H(const __grid_constant__ StencilParams p)
{
	PDL_PROLOGUE();
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int j = blockIdx.y * blockDim.y + threadIdx.y;
	const int k = p.k0 + blockIdx.z;
	// last kernel of the timestep: one thread advances numTS (no kernel of this launch reads it; saves the k_tick launch)
	if (p.tick && (blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x | threadIdx.y) == 0) *p.tick += 1;
	// UpdateCurrents stops one line short in every direction (engine.cpp:179-183; k1 <= held planes - 1: host)
	if (i >= p.nx - 1 || j >= p.ny - 1 || k >= p.k1) return;
	const long long o = (long long)k * p.plane + (long long)j * p.pitch + i;
	const float* __restrict__ V0 = p.V;
	const float* __restrict__ V1 = p.V + p.comp;
	const float* __restrict__ V2 = p.V + 2 * p.comp;
	const unsigned e = reinterpret_cast<const IdxT*>(p.idx)[o];
	const float v0c = V0[o], v1c = V1[o], v2c = V2[o];
	const float v0jp = V0[o + p.pitch], v2jp = V2[o + p.pitch];
	const float v0n = V0[o + p.plane], v1n = V1[o + p.plane];
	const float v1xp = V1[o + 1], v2xp = V2[o + 1];
	float c0 = p.I[o], c1 = p.I[p.comp + o], c2 = p.I[2 * p.comp + o];
	float* O = p.Iout ? p.Iout : p.I;   // out-of-place: the slab's top plane in the one-pass schedule (cells that are never updated hold the same value in both sets)
	const float4 A = __ldg(p.tA + e), B = __ldg(p.tB + e);
	const float curl0 = fadd(fsub(fsub(v2c, v2jp), v1c), v1n);
	const float curl1 = fadd(fsub(fsub(v0c, v0n), v2c), v2xp);
	const float curl2 = fadd(fsub(fsub(v1c, v1xp), v0c), v0jp);
	if (HAS_PML && A.w != 0.0f) {
		long long cs;
		const long long fo = pml_flux_offset(p, i, j, k, cs);
		if (fo < 0) { if (p.Iout) { O[o] = c0; O[p.comp + o] = c1; O[2 * p.comp + o] = c2; } return; }
		const float4 P0 = __ldg(p.tP0 + e), P1 = __ldg(p.tP1 + e), P2 = __ldg(p.tP2 + e);
		c0 = leap_pml(c0, A.x, B.x, curl0, P0.x, P1.x, P2.x, p.flux + fo, p.flux + fo);
		c1 = leap_pml(c1, A.y, B.y, curl1, P0.y, P1.y, P2.y, p.flux + fo + cs, p.flux + fo + cs);
		c2 = leap_pml(c2, A.z, B.z, curl2, P0.z, P1.z, P2.z, p.flux + fo + 2 * cs, p.flux + fo + 2 * cs);
	} else {
		c0 = leap(c0, A.x, B.x, curl0);
		c1 = leap(c1, A.y, B.y, curl1);
		c2 = leap(c2, A.z, B.z, curl2);
	}
	O[o] = c0; O[p.comp + o] = c1; O[2 * p.comp + o] = c2;
}

// ---------------------------------------------------------------------------------------
// UPML cells the stencil kernels do not visit.  Engine_Ext_UPML runs its pre/post hooks on
// EVERY cell of a box (engine_ext_upml.cpp:63-90), but UpdateCurrents skips the last line of
// each direction (engine.cpp:179-183).  For those cells pre+post collapse to
//   f = a_vv*I - a_fo*F ; F' = F ; I' = f + a_fn*F
// (the flux is swapped in and straight back out).  One thread per listed cell (list built at upload).
// ---------------------------------------------------------------------------------------
struct PmlEdgeParams {
	float* X;                 // I base
	const void* idx;
	const float4 *tP0, *tP1, *tP2;
	float* flux;
	const long long* cell;    // [count] cell offset (without component)
	const long long* fluxoff; // [count] flux offset of component 0
	const long long* fluxcs;  // [count] flux component stride of the entry's box
	long long count;
	long long comp;
};

template <typename IdxT>
__global__ void k_upml_untouched_H(const __grid_constant__ PmlEdgeParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	const long long o = p.cell[t];
	const unsigned e = reinterpret_cast<const IdxT*>(p.idx)[o];
	const float4 P0 = __ldg(p.tP0 + e), P1 = __ldg(p.tP1 + e), P2 = __ldg(p.tP2 + e);
	const long long cs = p.fluxcs[t], fo = p.fluxoff[t];
	const float a_vv[3] = {P0.x, P0.y, P0.z}, a_fn[3] = {P1.x, P1.y, P1.z}, a_fo[3] = {P2.x, P2.y, P2.z};
#pragma unroll
	for (int n = 0; n < 3; ++n) {
		const float F = p.flux[fo + n * cs];
		const float X = p.X[n * p.comp + o];
		const float f = fsub(fmul(a_vv[n], X), fmul(a_fo[n], F));
		p.X[n * p.comp + o] = fadd(f, fmul(a_fn[n], F));
	}
}

// ---------------------------------------------------------------------------------------
// Mur ABC: Engine_Ext_Mur_ABC engine_ext_mur_abc.cpp:82-173, all planes in one launch.
// ---------------------------------------------------------------------------------------
struct MurPlane {
	int ny, nyP, nyPP;
	int line, shift;
	int n0, n1;
	unsigned start_ts;
	long long eoff; // first entry of this plane in the concatenated arrays
};
struct MurParams {
	float* V;
	const float* cP; const float* cPP;
	float* vP; float* vPP;
	const unsigned* ovr_start; // [2][total]: start timestep of the earlier-inserted plane whose write to the same edge wins (0xffffffff: none)
	const unsigned* numTS;
	int nplanes;
	long long total;
	int pitch; long long plane, comp;
	int z0, zown0, zown1;
	MurPlane pl[OEMS_MAX_MUR];
};

__device__ __forceinline__ bool mur_locate(const MurParams& p, long long t, int& m, long long offs[2], long long offs_shift[2])
{
	if (t >= p.total) return false;
	m = 0;
	while (m + 1 < p.nplanes && t >= p.pl[m + 1].eoff) ++m;
	const MurPlane& M = p.pl[m];
	if (*p.numTS < M.start_ts) return false; // IsActive(), engine_ext_mur_abc.h:56
	const long long l = t - M.eoff;
	int pos[3], ps[3];
	pos[M.ny] = M.line; ps[M.ny] = M.shift;
	pos[M.nyP] = ps[M.nyP] = (int)(l / M.n1);
	pos[M.nyPP] = ps[M.nyPP] = (int)(l % M.n1);
	// ownership by the z of the boundary cell (slab sharding); the shift cell of a z-normal
	// plane is one plane further in and always held (ghost or owned)
	if (pos[2] < p.zown0 || pos[2] >= p.zown1) return false;
	const long long c = (long long)(pos[2] - p.z0) * p.plane + (long long)pos[1] * p.pitch + pos[0];
	const long long cs = (long long)(ps[2] - p.z0) * p.plane + (long long)ps[1] * p.pitch + ps[0];
	offs[0] = M.nyP * p.comp + c; offs[1] = M.nyPP * p.comp + c;
	offs_shift[0] = M.nyP * p.comp + cs; offs_shift[1] = M.nyPP * p.comp + cs;
	return true;
}

__global__ void k_mur_pre(const __grid_constant__ MurParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int m; long long o[2], os[2];
	if (!mur_locate(p, t, m, o, os)) return;
	p.vP[t] = fsub(p.V[os[0]], fmul(p.cP[t], p.V[o[0]]));
	p.vPP[t] = fsub(p.V[os[1]], fmul(p.cPP[t], p.V[o[1]]));
}
__global__ void k_mur_post(const __grid_constant__ MurParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int m; long long o[2], os[2];
	if (!mur_locate(p, t, m, o, os)) return;
	p.vP[t] = fadd(p.vP[t], fmul(p.cP[t], p.V[os[0]]));
	p.vPP[t] = fadd(p.vPP[t], fmul(p.cPP[t], p.V[os[1]]));
}
__device__ __forceinline__ void mur_apply_entry(const MurParams& p, long long t)
{
	int m; long long o[2], os[2];
	if (!mur_locate(p, t, m, o, os)) return;
	const unsigned ts = *p.numTS;
	if (ts < p.ovr_start[t]) p.V[o[0]] = p.vP[t];
	if (ts < p.ovr_start[p.total + t]) p.V[o[1]] = p.vPP[t];
}
__global__ void k_mur_apply(const __grid_constant__ MurParams p)
{
	PDL_PROLOGUE();
	mur_apply_entry(p, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}

// ---------------------------------------------------------------------------------------
// Excitation: Engine_Ext_Excitation::Apply2Voltages/Apply2Current engine_ext_excitation.cpp:33-94.
// Entries that hit the same (component, cell) are grouped at upload; one thread per group adds
// its entries in list order, so the fp32 result equals the sequential CPU loop.
// ---------------------------------------------------------------------------------------
struct ExcParams {
	float* X;
	const long long* tgt;      // [groups] field offset incl. component
	const unsigned* gstart;    // [groups+1]
	const float* amp;
	const unsigned* delay;
	const float* sig;
	const unsigned* numTS;
	unsigned groups, length, period;
};
__device__ __forceinline__ void excite_group(const ExcParams& p, unsigned g);
__global__ void k_excite(const __grid_constant__ ExcParams p)
{
	PDL_PROLOGUE();
	excite_group(p, blockIdx.x * blockDim.x + threadIdx.x);
}
// Apply2Voltages of the Mur planes and of the excitation in ONE launch (small meshes: a launch less per timestep).
// The reference runs Mur first, then the excitation (engine.cpp:239-244 over the priority-sorted list); the two write
// different cells -- checked at upload, Engine::mur_exc_disjoint -- so they may run side by side.
__global__ void k_mur_apply_excite(const __grid_constant__ MurParams m, const __grid_constant__ ExcParams e, unsigned mur_blocks)
{
	PDL_PROLOGUE();
	if (blockIdx.x < mur_blocks) mur_apply_entry(m, (long long)blockIdx.x * blockDim.x + threadIdx.x);
	else excite_group(e, (blockIdx.x - mur_blocks) * blockDim.x + threadIdx.x);
}
__device__ __forceinline__ void excite_group(const ExcParams& p, unsigned g)
{
	if (g >= p.groups) return;
	const int numTS = (int)*p.numTS;
	int per = numTS + 1;
	if (p.period > 0) per = (int)p.period;
	float x = p.X[p.tgt[g]];
	for (unsigned n = p.gstart[g]; n < p.gstart[g + 1]; ++n) {
		int pos = numTS - (int)p.delay[n];
		pos *= (pos > 0);
		pos %= per;
		pos *= (pos < (int)p.length);
		x = fadd(x, fmul(p.amp[n], p.sig[pos]));
	}
	p.X[p.tgt[g]] = x;
}

// ---------------------------------------------------------------------------------------
// Lorentz / Drude / conducting sheet ADE, one dispersion order per launch:
// pre-hooks engine_ext_lorentzmaterial.cpp:79-168, apply engine_ext_dispersive.cpp:76-127.
// ---------------------------------------------------------------------------------------
struct LorParams {
	float* X;               // V or I base
	const long long* cell;  // [count] cell offset (without component)
	const float* c_int;     // [3][count]
	const float* c_ext;
	const float* c_lor;     // may be NULL
	float* ade;             // [3][count]
	float* lor_ade;         // [3][count] or NULL
	unsigned count;
	long long comp;
	unsigned first;         // k_lorentz_apply: entries [first, count) only (the slab's top plane in the one-pass schedule)
	const unsigned short* pidx; // [count] index into ptab, or NULL (more than 65536 distinct tuples: per-cell arrays)
	const float* ptab;          // [n][9]: c_int[3], c_ext[3], c_lor[3]
};
__global__ void k_lorentz_pre(const __grid_constant__ LorParams p)
{
	PDL_PROLOGUE();
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= p.count) return;
	const long long c = p.cell[i];
#pragma unroll
	for (int n = 0; n < 3; ++n) {
		const size_t q = (size_t)n * p.count + i;
		const float x = p.X[n * p.comp + c];
		float a = p.ade[q];
		// compressed coefficients: a table of the distinct {int, ext, lor} x 3 tuples of the list + a 16-bit index per cell
		const float* t = p.pidx ? p.ptab + 9 * (size_t)p.pidx[i] : nullptr;
		const float ci = t ? __ldg(t + n) : p.c_int[q], ce = t ? __ldg(t + 3 + n) : p.c_ext[q];
		if (p.c_lor) {
			const float l = fadd(p.lor_ade[q], fmul(t ? __ldg(t + 6 + n) : p.c_lor[q], a));
			p.lor_ade[q] = l;
			a = fmul(a, ci);
			a = fadd(a, fmul(ce, fsub(x, l)));
		} else {
			a = fmul(a, ci);
			a = fadd(a, fmul(ce, x));
		}
		p.ade[q] = a;
	}
}
__global__ void k_lorentz_apply(const __grid_constant__ LorParams p)
{
	PDL_PROLOGUE();
	const unsigned i = p.first + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= p.count) return;
	const long long c = p.cell[i];
#pragma unroll
	for (int n = 0; n < 3; ++n) {
		const size_t q = (size_t)n * p.count + i;
		p.X[n * p.comp + c] = fsub(p.X[n * p.comp + c], p.ade[q]);
	}
}

// ---------------------------------------------------------------------------------------
// Lumped RLC: Engine_Ext_LumpedRLC engine_ext_lumpedRLC.cpp:83-142.  The reference rotates
// three array pointers; here the ring position is derived from numTS.
// ---------------------------------------------------------------------------------------
struct RlcParams {
	float* V;
	const long long* tgt; // field offset incl. component
	const float *ilv, *i2v, *vvd, *vv2, *vj1, *vj2, *ib0, *b1, *b2;
	float* Vd;  // [3][count] ring
	float* J;   // [3][count] ring
	float* Il;
	const unsigned* numTS;
	unsigned count;
};
// ring slot of logical index q (0 = newest) at timestep ts: after ts+1 rotations logical 0
// sits at physical (3 - (ts+1)%3)%3
__device__ __forceinline__ unsigned ring(unsigned ts_plus1, unsigned q) { return (q + 3u - ts_plus1 % 3u) % 3u; }
__global__ void k_rlc_pre(const __grid_constant__ RlcParams p)
{
	PDL_PROLOGUE();
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= p.count) return;
	const unsigned r = *p.numTS + 1; // rotation count after this call
	const float vd1 = p.Vd[(size_t)ring(r, 1) * p.count + i];
	p.Il[i] = fadd(p.Il[i], fmul(fmul(p.i2v[i], p.ilv[i]), vd1));
}
__global__ void k_rlc_apply(const __grid_constant__ RlcParams p)
{
	PDL_PROLOGUE();
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= p.count) return;
	const unsigned r = *p.numTS + 1;
	const size_t s0 = (size_t)ring(r, 0) * p.count + i, s1 = (size_t)ring(r, 1) * p.count + i,
	             s2 = (size_t)ring(r, 2) * p.count + i;
	const float v = p.V[p.tgt[i]];
	// vvd*(V - Il + vv2*Vd[2] + vj1*J[1] + vj2*J[2]), left to right
	float t = fsub(v, p.Il[i]);
	t = fadd(t, fmul(p.vv2[i], p.Vd[s2]));
	t = fadd(t, fmul(p.vj1[i], p.J[s1]));
	t = fadd(t, fmul(p.vj2[i], p.J[s2]));
	const float vd0 = fmul(p.vvd[i], t);
	p.Vd[s0] = vd0;
	float jn = fmul(p.ib0[i], fsub(vd0, p.Vd[s2]));
	jn = fsub(jn, fmul(fmul(p.b1[i], p.ib0[i]), p.J[s1]));
	jn = fsub(jn, fmul(fmul(p.b2[i], p.ib0[i]), p.J[s2]));
	p.J[s0] = jn;
	p.V[p.tgt[i]] = vd0;
}

// ---------------------------------------------------------------------------------------
// TFSF plane wave: Engine_Ext_TFSF::DoPostVoltageUpdates / DoPostCurrentUpdates
// (engine_ext_tfsf.cpp:36-215).  The face loops of the reference are flattened at upload into one
// ordered update list per field; updates of the same edge (box edges belong to two faces) are
// grouped and applied by one thread in the reference's order.  Per update:
//   X = float( X + (1.0 - dd)*amp*sig[lookup[d]] + dd*amp*sig[lookup[d+1]] )
// with the C++ types of the reference: the first product chain in double, the second in float.
// ---------------------------------------------------------------------------------------
struct TfsfParams {
	float* X;
	const long long* tgt;      // [groups] field offset incl. component
	const unsigned* gstart;    // [groups+1]
	const unsigned* delay;     // m_VoltDelay / m_CurrDelay
	const float* dd;           // m_VoltDelayDelta / m_CurrDelayDelta
	const float* amp;          // m_VoltAmp / m_CurrAmp
	const float* sig;          // current signal for the voltage update and vice versa
	const unsigned* numTS;
	unsigned groups, length, period;
};
__device__ __forceinline__ unsigned tfsf_lookup(unsigned numTS, unsigned n, unsigned length, unsigned p)
{
	unsigned v;
	if (numTS < n) v = 0;                                  // engine_ext_tfsf.cpp:45-52
	else if (numTS - n >= length && p == 0) v = 0;
	else v = numTS - n;
	if (p > 0) v %= p;
	return v;
}
__global__ void k_tfsf(const __grid_constant__ TfsfParams p)
{
	PDL_PROLOGUE();
	const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= p.groups) return;
	const unsigned numTS = *p.numTS;
	float x = p.X[p.tgt[g]];
	for (unsigned n = p.gstart[g]; n < p.gstart[g + 1]; ++n) {
		const unsigned d = p.delay[n];
		const float dd = p.dd[n], amp = p.amp[n];
		const float s1 = p.sig[tfsf_lookup(numTS, d, p.length, p.period)];
		const float s2 = p.sig[tfsf_lookup(numTS, d + 1, p.length, p.period)];
		double acc = __dadd_rn((double)x, __dmul_rn(__dmul_rn(__dsub_rn(1.0, (double)dd), (double)amp), (double)s1));
		acc = __dadd_rn(acc, (double)fmul(fmul(dd, amp), s2));
		x = (float)acc;
	}
	p.X[p.tgt[g]] = x;
}

// ---------------------------------------------------------------------------------------
// Local absorbing sheets: Engine_Ext_Absorbing_BC engine_ext_absorbing_bc.cpp:108-366 (1st order
// Mur on a sheet inside the mesh, optionally with super-absorption on H).  One list entry per
// sheet point and tangential component; pre/post/apply exactly as the reference's six hooks.
// ---------------------------------------------------------------------------------------
struct SheetParams {
	float* X;              // V or I of the set the hook works on
	const long long* o;    // [count] offset of the sheet value (incl. component)
	const long long* os;   // [count] offset of the shifted value
	const float* K1;       // [count]
	const float* K2;       // [count] (current lists of super-absorbing sheets)
	float* store;          // [count] m_V_nyP/nyPP or m_I_nyP/nyPP
	long long count;
};
__global__ void k_sheet_pre(const __grid_constant__ SheetParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	p.store[t] = fsub(p.X[p.os[t]], fmul(p.K1[t], p.X[p.o[t]]));          // :128-129, :246-247
}
__global__ void k_sheet_post(const __grid_constant__ SheetParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	p.store[t] = fadd(p.store[t], fmul(p.K1[t], p.X[p.os[t]]));           // :164-165, :303-304
}
__global__ void k_sheet_apply_V(const __grid_constant__ SheetParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	p.X[p.o[t]] = p.store[t];                                               // :199-200
}
__global__ void k_sheet_apply_I(const __grid_constant__ SheetParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.count) return;
	// (Hsa*K2 + Hc)/(K2 + 1.0): float numerator, double denominator                       :355-356
	const float num = fadd(fmul(p.store[t], p.K2[t]), p.X[p.o[t]]);
	p.X[p.o[t]] = (float)__ddiv_rn((double)num, __dadd_rn((double)p.K2[t], 1.0));
}

__global__ void k_tick(unsigned* numTS) {
	PDL_PROLOGUE(); *numTS += 1; }

// upload helper: expands an operator index given as unique xy planes + one plane id per z into the per-cell
// index idx[z][y][pitch] (padding cells point at the all-zero entry `pad`)
template <typename IdxT>
__global__ void k_expand_planes(IdxT* idx, const IdxT* uplanes, const unsigned* plane_of_z, int z0, long long rows, int nx, int ny, int pitch, IdxT pad)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; // one thread per (row, x)
	if (t >= rows * pitch) return;
	const long long r = t / pitch;
	const int x = (int)(t % pitch);
	const int zl = (int)(r / ny), y = (int)(r % ny);
	idx[t] = x < nx ? uplanes[((size_t)plane_of_z[z0 + zl] * ny + y) * nx + x] : pad;
}

// upload helper: operator index of the padding cells (i >= nx) points at the all-zero entry
template <typename IdxT>
__global__ void k_fill_index_padding(IdxT* idx, long long rows, int nx, int pitch, IdxT value)
{
	PDL_PROLOGUE();
	const int pad = pitch - nx;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (pad <= 0 || t >= rows * pad) return;
	idx[(t / pad) * pitch + nx + (int)(t % pad)] = value;
}

// ---------------------------------------------------------------------------------------
// Probes: ordered signed sums.  One warp per probe: lanes gather 32 terms at a time, then every
// lane accumulates them in list order through shuffles, so the result equals the reference's
// sequential loop (fp64 for voltages engine_interface_fdtd.cpp:206-232, fp32 for currents
// processcurrent.cpp:96-171).
// ---------------------------------------------------------------------------------------
struct ProbeParams {
	const float* V; const float* I;
	const long long* term_off;   // offset incl. component
	const signed char* term_sign;
	const unsigned* pstart;      // [nprobes+1] term ranges
	const unsigned* pvalue;      // [nprobes] first value slot
	const unsigned char* pkind;  // 0 voltage fp64, 1 current fp32, 2 raw E (3 values), 3 raw H
	double* out;                 // [nvalues] (+ slot offset by the caller)
	unsigned nprobes;
};
__global__ void k_probes(const __grid_constant__ ProbeParams p)
{
	PDL_PROLOGUE();
	const unsigned pr = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const unsigned lane = threadIdx.x % 32;
	if (pr >= p.nprobes) return;
	const unsigned t0 = p.pstart[pr], t1 = p.pstart[pr + 1];
	const unsigned kind = p.pkind[pr];
	double* out = p.out + p.pvalue[pr];
	if (kind >= 2) {
		if (lane < 3 && t0 + lane < t1) {
			const float* F = kind == 2 ? p.V : p.I;
			out[lane] = (double)F[p.term_off[t0 + lane]];
		}
		return;
	}
	const float* F = kind == 0 ? p.V : p.I;
	double acc64 = 0.0;
	float acc32 = 0.0f;
	for (unsigned base = t0; base < t1; base += 32) {
		float v = 0.0f;
		if (base + lane < t1) v = (float)p.term_sign[base + lane] * F[p.term_off[base + lane]];
		const unsigned cnt = min(32u, t1 - base);
		for (unsigned q = 0; q < cnt; ++q) {
			const float x = __shfl_sync(0xffffffffu, v, q);
			if (kind == 0) acc64 = __dadd_rn(acc64, (double)x);
			else acc32 = fadd(acc32, x);
		}
	}
	if (lane == 0) out[0] = kind == 0 ? acc64 : (double)acc32;
}

// ---------------------------------------------------------------------------------------
// Energy estimate: Engine_Interface_FDTD::CalcFastEnergy engine_interface_fdtd.cpp:302-347,
// fp32 products accumulated in fp64, warp-shuffle + block reduction, one atomic per block.
// ---------------------------------------------------------------------------------------
struct EnergyParams {
	const float* V; const float* I;
	int nx, ny;          // loop bounds are nx-1, ny-1
	int k0, k1;          // local plane range (owned, global k < nz-1)
	int pitch; long long plane, comp;
	double* acc;         // [2]: sum V^2, sum I^2
};
__global__ void k_energy(const __grid_constant__ EnergyParams p)
{
	PDL_PROLOGUE();
	double e = 0.0, h = 0.0;
	const long long rows = (long long)(p.k1 - p.k0) * (p.ny - 1);
	for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y) {
		const int k = p.k0 + (int)(r / (p.ny - 1));
		const int j = (int)(r % (p.ny - 1));
		const long long o = (long long)k * p.plane + (long long)j * p.pitch;
		for (int i = threadIdx.x; i < p.nx - 1; i += blockDim.x)
#pragma unroll
			for (int n = 0; n < 3; ++n) {
				const float v = p.V[n * p.comp + o + i], c = p.I[n * p.comp + o + i];
				e += (double)fmul(v, v);
				h += (double)fmul(c, c);
			}
	}
	for (int s = 16; s > 0; s >>= 1) {
		e += __shfl_down_sync(0xffffffffu, e, s);
		h += __shfl_down_sync(0xffffffffu, h, s);
	}
	__shared__ double se[32], sh[32];
	const int w = threadIdx.y;
	if (threadIdx.x == 0) { se[w] = e; sh[w] = h; }
	__syncthreads();
	if (threadIdx.y == 0 && threadIdx.x == 0) {
		double E = 0, H = 0;
		for (int q = 0; q < (int)blockDim.y; ++q) { E += se[q]; H += sh[q]; }
		atomicAdd(p.acc, E);
		atomicAdd(p.acc + 1, H);
	}
}

// ---------------------------------------------------------------------------------------
// Deterministic pre-fill and bit digest of the fields (measurement aids): bench.py starts its timed
// window from a filled domain instead of the all-zero one a point source leaves after a few steps, and prints a
// digest of E and H afterwards.  Both are functions of the GLOBAL cell index only, so z-slab engines fill the same
// values (ghost planes included) and the sum of the per-slab digests equals the single-GPU digest.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
	return x;
}
struct FillParams {
	float* V; float* I;
	int nx, ny, nzg;      // global mesh
	int z0, nzl;          // first held global plane, held planes
	int pitch; long long plane, comp;
	unsigned long long seed;
};
// value = ((hash(i,j,k,n,field) mod 2^16) - 2^15) * 1e-6 (the recipe of SURVEY 8d); the last line of every
// direction keeps H = 0 as in any reachable state (ii = iv = 0 there, operator.cpp:1176-1183)
__global__ void k_fill(const __grid_constant__ FillParams p)
{
	PDL_PROLOGUE();
	const long long rows = (long long)p.nzl * p.ny;
	for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y) {
		const int kl = (int)(r / p.ny), j = (int)(r % p.ny);
		const int k = p.z0 + kl;
		const long long o = (long long)kl * p.plane + (long long)j * p.pitch;
		for (int i = threadIdx.x; i < p.nx; i += blockDim.x) {
			const unsigned long long g = ((unsigned long long)k * p.ny + j) * p.nx + i;
#pragma unroll
			for (int n = 0; n < 3; ++n) {
				const unsigned long long hv = mix64(g * 6 + n + p.seed * 0x9E3779B97F4A7C15ull);
				const unsigned long long hi = mix64(g * 6 + 3 + n + p.seed * 0x9E3779B97F4A7C15ull);
				p.V[n * p.comp + o + i] = (float)((int)(hv & 0xffff) - 32768) * 1e-6f;
				const bool last = (i == p.nx - 1) || (j == p.ny - 1) || (k == p.nzg - 1);
				p.I[n * p.comp + o + i] = last ? 0.f : (float)((int)(hi & 0xffff) - 32768) * 2.6e-9f;
			}
		}
	}
}
struct DigestParams {
	const float* X;
	int nx, ny;
	int z0, k0, k1;       // first held global plane; local plane range of the OWNED planes
	int pitch; long long plane, comp;
	unsigned long long* acc;
};
// order-independent: sum over owned cells and components of mix(global index, bits) mod 2^64
__global__ void k_digest(const __grid_constant__ DigestParams p)
{
	PDL_PROLOGUE();
	unsigned long long d = 0;
	const long long rows = (long long)(p.k1 - p.k0) * p.ny;
	for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y) {
		const int kl = p.k0 + (int)(r / p.ny), j = (int)(r % p.ny);
		const long long o = (long long)kl * p.plane + (long long)j * p.pitch;
		for (int i = threadIdx.x; i < p.nx; i += blockDim.x) {
			const unsigned long long g = ((unsigned long long)(p.z0 + kl) * p.ny + j) * p.nx + i;
#pragma unroll
			for (int n = 0; n < 3; ++n) {
				unsigned bits = __float_as_uint(p.X[n * p.comp + o + i]);
				if (bits == 0x80000000u) bits = 0; // -0 and +0 are the same field value
				d += mix64((g * 3 + n) * 0x100000001b3ull ^ ((unsigned long long)bits << 17) ^ bits);
			}
		}
	}
	for (int s = 16; s > 0; s >>= 1) d += __shfl_down_sync(0xffffffffu, d, s);
	if (threadIdx.x == 0) atomicAdd(p.acc, d);
}

// ---------------------------------------------------------------------------------------
// Steady-state detection: Engine_Ext_SteadyState::Apply2Voltages engine_ext_steadystate.cpp:50-107.
// Every timestep the probe voltages go into a 2-period ring; when a period completes
// (TS % p == 0, TS >= 2p) the energy estimate of that instant (E after the stencil, H before its
// update) is taken and ring + energies are snapshotted for the host, which evaluates the
// criterion without per-step read-backs.
// ---------------------------------------------------------------------------------------
struct SsParams {
	const float* V; const float* I;
	const long long* off;   // [count] voltage offsets incl. component
	double* rec;            // ring [2p][count]
	double* snap;           // snapshot of the ring at the last completed period
	double* energy;         // [0],[1] scratch sums V^2, I^2 ; [2] previous, [3] current period energy
	unsigned* info;         // [0] number of completed checks, [1] TS of the last one
	const unsigned* numTS;
	unsigned period, count;
	int nx, ny, k0, k1;
	int pitch; long long plane, comp;
};
__device__ __forceinline__ bool ss_due(const SsParams& p) { const unsigned ts = *p.numTS; return ts % p.period == 0 && ts >= 2 * p.period; }
__global__ void k_ss_record(const __grid_constant__ SsParams p)
{
	PDL_PROLOGUE();
	const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned ts = *p.numTS;
	if (n < p.count) p.rec[(size_t)(ts % (2 * p.period)) * p.count + n] = (double)p.V[p.off[n]];
	if (n == 0 && ss_due(p)) { p.energy[0] = 0.0; p.energy[1] = 0.0; }
}
__global__ void k_ss_energy(const __grid_constant__ SsParams p)
{
	PDL_PROLOGUE();
	if (!ss_due(p)) return;
	double e = 0.0, h = 0.0;
	const long long rows = (long long)(p.k1 - p.k0) * (p.ny - 1);
	for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y) {
		const int k = p.k0 + (int)(r / (p.ny - 1));
		const int j = (int)(r % (p.ny - 1));
		const long long o = (long long)k * p.plane + (long long)j * p.pitch;
		for (int i = threadIdx.x; i < p.nx - 1; i += blockDim.x)
#pragma unroll
			for (int n = 0; n < 3; ++n) {
				const float v = p.V[n * p.comp + o + i], c = p.I[n * p.comp + o + i];
				e += (double)fmul(v, v);
				h += (double)fmul(c, c);
			}
	}
	for (int s = 16; s > 0; s >>= 1) {
		e += __shfl_down_sync(0xffffffffu, e, s);
		h += __shfl_down_sync(0xffffffffu, h, s);
	}
	if (threadIdx.x == 0) { atomicAdd(p.energy, e); atomicAdd(p.energy + 1, h); }
}
__global__ void k_ss_snapshot(const __grid_constant__ SsParams p)
{
	PDL_PROLOGUE();
	if (!ss_due(p)) return;
	const size_t n = (size_t)2 * p.period * p.count;
	for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) p.snap[q] = p.rec[q];
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		p.energy[2] = p.energy[3];
		p.energy[3] = 8.85418781762e-12 * p.energy[0] + 1.256637062e-6 * p.energy[1];
		p.info[0] += 1;
		p.info[1] = *p.numTS;
	}
}

// ---------------------------------------------------------------------------------------
// Field dump: ProcessFields::CalcField processfields.cpp:283-409 + interpolation
// engine_interface_fdtd.cpp:63-124 (E) / :150-204 (H), evaluated in fp64 like the reference
// and stored as fp32 in {3,nz,ny,nx} order, x fastest.
// ---------------------------------------------------------------------------------------
struct DumpParams {
	const float* V; const float* I;
	const unsigned *px, *py, *pz;
	const double* el[3];   // primal edge length per line
	const double* del[3];  // dual edge length per line
	float* out;
	int is_H, interp;
	unsigned onx, ony, onz;
	int nx, ny, gnz, z0;
	int pitch; long long plane, comp;
};
// ---------------------------------------------------------------------------------------
// ProcessFieldsFD::Process (Common/processfields_fd.cpp:72-107): running DFT of a field dump,
//   field_fd[f] += field_td * exp_jwt_2_dt[f]      (complex<float> += float * complex<float>)
// with the weights computed by the caller exactly as the reference does.  The time-domain
// sample is the output of k_dump, which stays on the device.
// ---------------------------------------------------------------------------------------
struct FdParams {
	const float* td;   // [3*count] time-domain dump
	float2* acc;       // [nfreq][3*count]
	const float2* w;   // [nfreq] weights of this sample
	long long n;       // 3*count
	unsigned nfreq;
};
__global__ void k_fd_accumulate(const __grid_constant__ FdParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= p.n) return;
	const float v = p.td[t];
	for (unsigned f = 0; f < p.nfreq; ++f) {
		const float2 w = p.w[f];
		float2 a = p.acc[(long long)f * p.n + t];
		a.x = fadd(a.x, fmul(v, w.x));
		a.y = fadd(a.y, fmul(v, w.y));
		p.acc[(long long)f * p.n + t] = a;
	}
}

__device__ __forceinline__ double dump_raw(const DumpParams& p, int is_H, int n, const int pos[3])
{
	const long long o = n * p.comp + (long long)(pos[2] - p.z0) * p.plane + (long long)pos[1] * p.pitch + pos[0];
	const double value = is_H ? (double)p.I[o] : (double)p.V[o];
	const double delta = is_H ? p.del[n][pos[n]] : p.el[n][pos[n]];
	if (delta != 0.0) return __ddiv_rn(value, delta);
	return 0.0;
}
// Engine_Interface_FDTD::GetEField / GetHField (engine_interface_fdtd.cpp:63-124,150-204): the
// field at mesh position pos with interpolation type p.interp (0 none, 1 node, 2 cell), fp64
__device__ __forceinline__ void field_interp(const DumpParams& p, const int pos[3], double out[3])
{
	const int N[3] = {p.nx, p.ny, p.gnz};
	int ip[3] = {pos[0], pos[1], pos[2]};
	if (!p.is_H) {
		if (p.interp == 1) {
			for (int n = 0; n < 3; ++n) {
				if (pos[n] == N[n] - 1) { --ip[n]; out[n] = dump_raw(p, 0, n, ip); ++ip[n]; continue; }
				const double delta = p.el[n][ip[n]];
				out[n] = dump_raw(p, 0, n, ip);
				if (delta == 0) { out[n] = 0; continue; }
				if (pos[n] == 0) continue;
				--ip[n];
				const double dDown = p.el[n][ip[n]];
				const double dRel = __ddiv_rn(delta, __dadd_rn(delta, dDown));
				out[n] = __dadd_rn(__dmul_rn(out[n], __dsub_rn(1.0, dRel)), __dmul_rn(dump_raw(p, 0, n, ip), dRel));
				++ip[n];
			}
		} else if (p.interp == 2) {
			for (int n = 0; n < 3; ++n) {
				const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
				if (pos[0] == N[0] - 1 || pos[1] == N[1] - 1 || pos[2] == N[2] - 1) { out[n] = 0; continue; }
				double a = dump_raw(p, 0, n, ip);
				++ip[nP]; a = __dadd_rn(a, dump_raw(p, 0, n, ip));
				++ip[nPP]; a = __dadd_rn(a, dump_raw(p, 0, n, ip));
				--ip[nP]; a = __dadd_rn(a, dump_raw(p, 0, n, ip));
				--ip[nPP];
				out[n] = __ddiv_rn(a, 4.0);
			}
		} else {
			for (int n = 0; n < 3; ++n) out[n] = dump_raw(p, 0, n, pos);
		}
	} else {
		if (p.interp == 1) {
			for (int n = 0; n < 3; ++n) {
				const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
				if (pos[0] == N[0] - 1 || pos[1] == N[1] - 1 || pos[2] == N[2] - 1 || pos[nP] == 0 || pos[nPP] == 0) { out[n] = 0; continue; }
				double a = dump_raw(p, 1, n, ip);
				--ip[nP]; a = __dadd_rn(a, dump_raw(p, 1, n, ip));
				--ip[nPP]; a = __dadd_rn(a, dump_raw(p, 1, n, ip));
				++ip[nP]; a = __dadd_rn(a, dump_raw(p, 1, n, ip));
				++ip[nPP];
				out[n] = __ddiv_rn(a, 4.0);
			}
		} else if (p.interp == 2) {
			for (int n = 0; n < 3; ++n) {
				const double delta = p.del[n][ip[n]];
				out[n] = dump_raw(p, 1, n, ip);
				if (pos[n] >= N[n] - 1) { out[n] = 0; continue; }
				++ip[n];
				const double dUp = p.del[n][ip[n]];
				const double dRel = __ddiv_rn(delta, __dadd_rn(delta, dUp));
				out[n] = __dadd_rn(__dmul_rn(out[n], __dsub_rn(1.0, dRel)), __dmul_rn(dump_raw(p, 1, n, ip), dRel));
				--ip[n];
			}
		} else {
			for (int n = 0; n < 3; ++n) out[n] = dump_raw(p, 1, n, pos);
		}
	}
}

__global__ void k_dump(const __grid_constant__ DumpParams p)
{
	PDL_PROLOGUE();
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long cnt = (long long)p.onx * p.ony * p.onz;
	if (t >= cnt) return;
	const unsigned ox = (unsigned)(t % p.onx), oy = (unsigned)((t / p.onx) % p.ony), oz = (unsigned)(t / ((long long)p.onx * p.ony));
	const int pos[3] = {(int)p.px[ox], (int)p.py[oy], (int)p.pz[oz]};
	double out[3];
	field_interp(p, pos, out);
	p.out[t] = (float)out[0];
	p.out[cnt + t] = (float)out[1];
	p.out[2 * cnt + t] = (float)out[2];
}

// ---------------------------------------------------------------------------------------
// ProcessModeMatch::CalcMultipleIntegrals (Common/processmodematch.cpp:222-266): projection of the
// node-interpolated tangential field on a plane onto a mode template,
//   value  += field_n * dist_n * area      purity += field_n * field_n * area     (n = 0, 1)
// summed in the reference's loop order (posP outer, posPP inner, n innermost): one warp gathers 32
// points at a time, then accumulates their terms in order through shuffles.
// ---------------------------------------------------------------------------------------
struct ModeParams {
	DumpParams d;          // fields, edge lengths, interpolation (node), mesh
	int ny;                // plane normal
	int line;              // start[ny]
	int startP, startPP;   // first line in nP = (ny+1)%3, nPP = (ny+2)%3
	unsigned nl0, nl1;     // m_numLines
	const double* dist0;   // [nl0*nl1] normalised mode template of component nP
	const double* dist1;   // component nPP
	const double* area;    // Op->GetNodeArea(ny, pos, dualMesh)
	double* out;           // [3]: value, value^2 / purity, purity (raw sum: z-slab partial results are added on the host)
	int own_z0, own_z1;    // only points with own_z0 <= z < own_z1 contribute (the planes a z-slab engine owns)
};
__global__ void k_mode_match(const __grid_constant__ ModeParams p)
{
	PDL_PROLOGUE();
	const unsigned lane = threadIdx.x;
	const unsigned npts = p.nl0 * p.nl1;
	const int nP = (p.ny + 1) % 3, nPP = (p.ny + 2) % 3;
	double value = 0.0, purity = 0.0;
	for (unsigned base = 0; base < npts; base += 32) {
		double tv0 = 0.0, tv1 = 0.0, tp0 = 0.0, tp1 = 0.0;
		const unsigned q = base + lane;
		if (q < npts) {
			int pos[3];
			pos[p.ny] = p.line;
			pos[nP] = p.startP + (int)(q / p.nl1);
			pos[nPP] = p.startPP + (int)(q % p.nl1);
			if (pos[2] >= p.own_z0 && pos[2] < p.own_z1) {
				double f[3];
				field_interp(p.d, pos, f);
				const double a = p.area[q], f0 = f[nP], f1 = f[nPP];
				tv0 = __dmul_rn(__dmul_rn(f0, p.dist0[q]), a); tp0 = __dmul_rn(__dmul_rn(f0, f0), a);
				tv1 = __dmul_rn(__dmul_rn(f1, p.dist1[q]), a); tp1 = __dmul_rn(__dmul_rn(f1, f1), a);
			}
		}
		const unsigned cnt = min(32u, npts - base);
		for (unsigned s = 0; s < cnt; ++s) {
			value = __dadd_rn(value, __shfl_sync(0xffffffffu, tv0, s));
			purity = __dadd_rn(purity, __shfl_sync(0xffffffffu, tp0, s));
			value = __dadd_rn(value, __shfl_sync(0xffffffffu, tv1, s));
			purity = __dadd_rn(purity, __shfl_sync(0xffffffffu, tp1, s));
		}
	}
	if (lane == 0) {
		p.out[0] = value;
		p.out[1] = purity != 0.0 ? __ddiv_rn(__dmul_rn(value, value), purity) : 0.0;
		p.out[2] = purity;
	}
}

// ---------------------------------------------------------------------------------------
// Multi-GPU halo: copy the two tangential components of one xy plane into the neighbour's
// ghost plane through an NVLink peer mapping, then publish the step number in the neighbour's
// flag.  Template: Engine_MPI::SendReceiveVoltages/Currents engine_mpi.cpp:84-182 (which
// packs into a host buffer and MPI_Isend/Irecv's it).
// ---------------------------------------------------------------------------------------
struct HaloParams {
	const float* src;       // local field base (V or I)
	float* dst;             // peer field base (mapped)
	long long src_plane_off, dst_plane_off; // k*plane offsets
	long long src_comp, dst_comp;
	long long n;            // floats per plane (pitch*ny), multiple of 4
	unsigned* done_counter; // local scratch
	volatile unsigned* peer_flag;
	const unsigned* numTS;
	unsigned flag_add;      // value published = numTS + flag_add
};
__global__ void k_halo_push(const __grid_constant__ HaloParams p)
{
	PDL_PROLOGUE();
	const long long n4 = p.n / 4;
	for (int c = 0; c < 2; ++c) {
		const float4* s = reinterpret_cast<const float4*>(p.src + c * p.src_comp + p.src_plane_off);
		float4* d = reinterpret_cast<float4*>(p.dst + c * p.dst_comp + p.dst_plane_off);
		for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x)
			d[q] = s[q];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned prev = atomicAdd(p.done_counter, 1u);
		if (prev == gridDim.x - 1) {
			*p.done_counter = 0;
			__threadfence_system();
			*p.peer_flag = *p.numTS + p.flag_add;
			__threadfence_system();
		}
	}
}
struct WaitParams {
	volatile unsigned* flag;
	const unsigned* numTS;
	unsigned flag_add;
	unsigned* error;
	long long timeout_cycles;
};
__global__ void k_halo_wait(const __grid_constant__ WaitParams p)
{
	PDL_PROLOGUE();
	const unsigned want = *p.numTS + p.flag_add;
	const long long t0 = clock64();
	while ((int)(*p.flag - want) < 0) {
		if (clock64() - t0 > p.timeout_cycles) {
			// the neighbour never published this step: stop here instead of stepping on with a stale ghost plane.
			// The trap makes every later call on this device fail (reported by iterate / sync), nothing is consumed.
			*p.error = 1;
			__threadfence_system();
			__trap();
		}
		__nanosleep(200);
	}
	__threadfence_system();
}

// ---------------------------------------------------------------------------------------
// Complete ghost planes for the readout (field dumps, FD dumps, mode matching on z-slab engines): the time
// loop only exchanges what the stencil needs (tangential E down, tangential H up).  The interpolating
// readers (engine_interface_fdtd.cpp:63-124,150-204) also take the normal components and E of the plane
// below / H of the plane above, so before a readout every slab pushes all three components of E AND H of its
// lowest owned plane into the lower neighbour's upper ghost plane and of its highest owned plane into the upper
// neighbour's lower ghost plane, then publishes the exchange number.  The values the time loop reads from the
// ghost planes are rewritten with the same bits; the parts it never reads (E below, H above, normal
// components) are free.  Template: the full-plane exchange Engine_MPI does every timestep
// (engine_mpi.cpp:84-182).
// ---------------------------------------------------------------------------------------
struct GhostPushParams {
	const float* srcV; const float* srcI;   // local field bases
	float* dstV; float* dstI;               // peer field bases (mapped)
	long long src_plane_off, dst_plane_off;
	long long src_comp, dst_comp;
	long long n;                            // floats per plane, multiple of 4
	unsigned* done_counter;
	volatile unsigned* peer_flag;
	unsigned value;
};
__global__ void k_ghost_push(const __grid_constant__ GhostPushParams p)
{
	PDL_PROLOGUE();
	const long long n4 = p.n / 4;
	for (int c = 0; c < 6; ++c) {
		const float* sb = c < 3 ? p.srcV : p.srcI;
		float* db = c < 3 ? p.dstV : p.dstI;
		const float4* s = reinterpret_cast<const float4*>(sb + (c % 3) * p.src_comp + p.src_plane_off);
		float4* d = reinterpret_cast<float4*>(db + (c % 3) * p.dst_comp + p.dst_plane_off);
		for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x)
			d[q] = s[q];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned prev = atomicAdd(p.done_counter, 1u);
		if (prev == gridDim.x - 1) {
			*p.done_counter = 0;
			__threadfence_system();
			*p.peer_flag = p.value;
			__threadfence_system();
		}
	}
}
// publish / wait for an exchange number (flags of the ghost exchange; the step flags use k_halo_wait)
struct FlagParams {
	volatile unsigned* flag[2];
	unsigned value;
	unsigned* error;
	long long timeout_cycles;
};
__global__ void k_flag_set(const __grid_constant__ FlagParams p)
{
	PDL_PROLOGUE();
	__threadfence_system();
	for (int q = 0; q < 2; ++q)
		if (p.flag[q]) *p.flag[q] = p.value;
	__threadfence_system();
}
__global__ void k_flag_wait(const __grid_constant__ FlagParams p)
{
	PDL_PROLOGUE();
	const long long t0 = clock64();
	for (int q = 0; q < 2; ++q) {
		if (!p.flag[q]) continue;
		while ((int)(*p.flag[q] - p.value) < 0) {
			if (clock64() - t0 > p.timeout_cycles) {
				*p.error = 1;
				__threadfence_system();
				__trap();
			}
			__nanosleep(200);
		}
	}
	__threadfence_system();
}
