// engine.cu -- host side of the B200 FDTD engine: upload, operator compression, extension
// schedule, CUDA-graph time stepping, readout.  No CPU compute fallback exists: every field
// update runs in the kernels of kernels.cuh.
#include "engine.h"

#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>

#define CK(call)                                                         \
	do {                                                                 \
		cudaError_t e_ = (call);                                         \
		if (e_ != cudaSuccess) return check(e_, #call);                  \
	} while (0)

int Engine::check(cudaError_t e, const char* what)
{
	if (e == cudaSuccess) return 0;
	err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
	return 1;
}

// Every kernel launch goes through here.  With option "pdl" = 1 the launch carries the programmatic-stream-
// serialization attribute (captured into the step graphs as programmatic edges): the next kernel of the stream may be
// set up while this one is still running; every kernel starts with PDL_PROLOGUE (kernels.cuh: griddepcontrol.wait), so
// it touches memory only after its predecessor has completed and flushed.  Measured (profiles/experiments_r02.md #8):
// no gain -- the 16-28 us timesteps of the small configs are the latency of the dependent loads inside the two
// stencil kernels, not launch gaps -- so it is off by default (bit-identical either way, tests run both).
static bool g_pdl = false;
template <typename... KArgs, typename... Args>
static void launch_k(void (*kern)(KArgs...), dim3 g, dim3 b, size_t smem, cudaStream_t s, Args&&... args)
{
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = s;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
	cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <typename T> T* Engine::dalloc(size_t n, bool zero)
{
	T* p = nullptr;
	if (n == 0) n = 1;
	if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
	if (zero) cudaMemsetAsync(p, 0, n * sizeof(T), stream);
	allocs.push_back(p);
	hbm_bytes += n * sizeof(T);
	return p;
}
template <typename T> T* Engine::upload(const std::vector<T>& v)
{
	T* p = dalloc<T>(v.size(), v.empty());
	if (p && !v.empty()) {
		cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
		h2d_bytes += v.size() * sizeof(T);
	}
	return p;
}

Engine::Engine(unsigned nx, unsigned ny, unsigned nz, int dev)
{
	gn[0] = nx; gn[1] = ny; gn[2] = nz;
	device = dev;
	zb = 0; ze = nz; z0 = 0; nzl = (int)nz;
	pitch = (int)((nx + 31) / 32 * 32);
	plane = (long long)pitch * ny;
	comp = plane * nzl;
}

Engine::~Engine()
{
	cudaSetDevice(device);
	if (stream) cudaStreamSynchronize(stream);
	if (graph_exec) cudaGraphExecDestroy(graph_exec);
	if (graph) cudaGraphDestroy(graph);
	for (int q = 0; q < 2; ++q) {
		if (graphf_exec[q]) cudaGraphExecDestroy(graphf_exec[q]);
		if (graphf[q]) cudaGraphDestroy(graphf[q]);
	}
	for (auto& d : dumps) {
		if (d.h_pinned) cudaFreeHost(d.h_pinned);
		if (d.ev_computed) cudaEventDestroy(d.ev_computed);
		if (d.ev_copied) cudaEventDestroy(d.ev_copied);
	}
	if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
	if (side_stream) {
		cudaStreamSynchronize(side_stream);
		cudaStreamDestroy(side_stream);
		for (int q = 0; q < 2; ++q) {
			if (ev_fork[q][0]) cudaEventDestroy(ev_fork[q][0]);
			if (ev_fork[q][1]) cudaEventDestroy(ev_fork[q][1]);
			if (ev_join[q]) cudaEventDestroy(ev_join[q]);
		}
	}
	for (auto& f : fds) {
		if (f.h_w) cudaFreeHost(f.h_w);
		for (int r = 0; r < FdHost::RING; ++r) if (f.ev[r]) cudaEventDestroy(f.ev[r]);
	}
	for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
	for (void* p : allocs) cudaFree(p);
	if (stream) cudaStreamDestroy(stream);
}

int Engine::set_slab(unsigned b, unsigned e)
{
	if (have_dense || have_compressed || finalized) return fail("set_slab must precede all uploads");
	if (b >= e || e > gn[2]) return fail("set_slab: bad range");
	zb = b; ze = e;
	z0 = (int)b - (b > 0 ? 1 : 0);
	const int z1 = (int)e + (e < gn[2] ? 1 : 0);
	nzl = z1 - z0;
	comp = plane * nzl;
	slab_set = true;
	return 0;
}

// ------------------------------------------------------------------------------ uploads
int Engine::set_operator_dense(const float* vv, const float* vi, const float* ii, const float* iv)
{
	if (finalized) return fail("engine already finalized");
	if (!vv || !vi || !ii || !iv) return fail("set_operator_dense: null pointer");
	const float* src[4] = {vv, vi, ii, iv};
	const size_t nl = (size_t)3 * gn[0] * gn[1] * nzl;
	for (int a = 0; a < 4; ++a) {
		h_dense[a].resize(nl);
		// keep ArrayNIJK order but only the held planes: [n][i][j][kl]
#pragma omp parallel for collapse(2) schedule(static)
		for (int n = 0; n < 3; ++n)
			for (unsigned i = 0; i < gn[0]; ++i)
				for (unsigned j = 0; j < gn[1]; ++j) {
					const float* s = src[a] + (((size_t)n * gn[0] + i) * gn[1] + j) * gn[2] + z0;
					float* d = h_dense[a].data() + (((size_t)n * gn[0] + i) * gn[1] + j) * nzl;
					memcpy(d, s, (size_t)nzl * sizeof(float));
				}
	}
	have_dense = true;
	have_compressed = false;
	return 0;
}

// development aid: OEMS_TIMING=1 prints how long the stages of the operator upload take
struct StageTimer {
	bool on; cudaStream_t st; double t0; const char* what;
	static double now() { return omp_get_wtime(); }
	StageTimer(const char* w, cudaStream_t s) : on(getenv("OEMS_TIMING") != nullptr), st(s), t0(0), what(w) { if (on) t0 = now(); }
	void lap(const char* name) { if (!on) return; cudaStreamSynchronize(st); const double t = now(); fprintf(stderr, "[oems timing] %s: %s %.1f ms\n", what, name, (t - t0) * 1e3); t0 = t; }
};

// operator index as unique xy planes + a plane id per z, expanded on the device
int Engine::set_operator_planes(unsigned nu, const oems_coeff_entry* table, unsigned n_uplanes, const void* uplanes,
                                const unsigned* plane_of_z, int ib)
{
	if (finalized) return fail("engine already finalized");
	if (!table || !uplanes || !plane_of_z || nu == 0 || n_uplanes == 0) return fail("set_operator_planes: null/empty input");
	if (ib != 2 && ib != 4) return fail("set_operator_planes: index_bytes must be 2 or 4");
	if (ib == 2 && nu > 65535) return fail("set_operator_planes: more than 65535 entries need a 32-bit index");
	for (unsigned k = 0; k < gn[2]; ++k)
		if (plane_of_z[k] >= n_uplanes) return fail("set_operator_planes: plane id out of range");
	CK(cudaSetDevice(device));
	if (!stream) CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
	if (d_idx) return fail("operator already uploaded");
	h_table.assign(table, table + nu);
	index_bytes = ib;
	n_unique = nu;
	const size_t cells = (size_t)plane * nzl, np = (size_t)gn[0] * gn[1];
	void* p = nullptr;
	CK(cudaMalloc(&p, cells * ib));
	allocs.push_back(p);
	hbm_bytes += cells * ib;
	d_idx = p;
	void* d_up = nullptr;
	unsigned* d_pz = nullptr;
	CK(cudaMalloc(&d_up, np * n_uplanes * ib));
	CK(cudaMalloc((void**)&d_pz, gn[2] * sizeof(unsigned)));
	CK(cudaMemcpyAsync(d_up, uplanes, np * n_uplanes * ib, cudaMemcpyHostToDevice, stream));
	CK(cudaMemcpyAsync(d_pz, plane_of_z, gn[2] * sizeof(unsigned), cudaMemcpyHostToDevice, stream));
	h2d_bytes += np * n_uplanes * ib + gn[2] * sizeof(unsigned);
	const long long rows = (long long)gn[1] * nzl, n = rows * pitch;
	const unsigned blocks = (unsigned)((n + 255) / 256);
	if (ib == 2) launch_k(k_expand_planes<uint16_t>, blocks, 256, 0, stream, (uint16_t*)p, (const uint16_t*)d_up, d_pz, z0, rows, (int)gn[0], (int)gn[1], pitch, (uint16_t)nu);
	else launch_k(k_expand_planes<uint32_t>, blocks, 256, 0, stream, (uint32_t*)p, (const uint32_t*)d_up, d_pz, z0, rows, (int)gn[0], (int)gn[1], pitch, (uint32_t)nu);
	CK(cudaStreamSynchronize(stream)); // the caller's buffers are free again
	CK(cudaGetLastError());
	cudaFree(d_up);
	cudaFree(d_pz);
	h2d_index_bytes = np * n_uplanes * ib + gn[2] * sizeof(unsigned);
	have_compressed = true;
	have_dense = false;
	return 0;
}

int Engine::set_operator_compressed(unsigned nu, const oems_coeff_entry* table, const void* index, int ib)
{
	if (finalized) return fail("engine already finalized");
	if (!table || !index || nu == 0) return fail("set_operator_compressed: null/empty input");
	if (ib != 2 && ib != 4) return fail("set_operator_compressed: index_bytes must be 2 or 4");
	if (ib == 2 && nu > 65535) return fail("set_operator_compressed: more than 65535 entries need a 32-bit index");
	CK(cudaSetDevice(device));
	if (!stream) CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
	StageTimer tmu("set_operator_compressed", nullptr);
	h_table.assign(table, table + nu);
	index_bytes = ib;
	n_unique = nu;
	// index goes straight to HBM, re-pitched; padding cells point at the zero entry nu
	const size_t cells = (size_t)plane * nzl;
	if (d_idx) return fail("operator already uploaded");
	void* p = nullptr;
	CK(cudaMalloc(&p, cells * ib));
	allocs.push_back(p);
	hbm_bytes += cells * ib;
	d_idx = p;
	{
		const long long rows = (long long)gn[1] * nzl;
		const long long padn = rows * (pitch - (int)gn[0]);
		if (padn > 0) {
			const unsigned blocks = (unsigned)((padn + 255) / 256);
			if (ib == 2) launch_k(k_fill_index_padding<uint16_t>, blocks, 256, 0, stream, (uint16_t*)p, rows, (int)gn[0], pitch, (uint16_t)nu);
			else launch_k(k_fill_index_padding<uint32_t>, blocks, 256, 0, stream, (uint32_t*)p, rows, (int)gn[0], pitch, (uint32_t)nu);
		}
	}
	const char* src = (const char*)index + (size_t)z0 * gn[1] * gn[0] * ib;
	cudaPointerAttributes attr;
	const bool src_pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	h2d_bytes += (size_t)gn[0] * ib * gn[1] * nzl;
	cudaGetLastError();
	if (src_pinned) {
		// page-locked source (oems_synth_pin, cudaHostRegister / cudaMallocHost by the caller): one DMA
		const size_t row_bytes = (size_t)gn[0] * ib;
		CK(cudaMemcpy2DAsync(p, (size_t)pitch * ib, src, row_bytes, row_bytes, (size_t)gn[1] * nzl, cudaMemcpyHostToDevice, stream));
		// not waited for here: the copy runs while finalize() allocates and clears the field sets; the caller
		// keeps the (page-locked) buffer alive until oems_cuda_finalize has returned (include/openems_b200.h)
	} else {
		// double-buffered pinned staging: the caller's buffer is pageable, a direct copy would run
		// at a fraction of the PCIe rate
		const size_t row_bytes = (size_t)gn[0] * ib, total_rows = (size_t)gn[1] * nzl;
		const size_t chunk_rows = std::max<size_t>(1, (size_t)(64u << 20) / row_bytes);
		void* stage[2] = {nullptr, nullptr};
		cudaEvent_t done[2];
		for (int q = 0; q < 2; ++q) {
			CK(cudaMallocHost(&stage[q], chunk_rows * row_bytes));
			CK(cudaEventCreateWithFlags(&done[q], cudaEventDisableTiming));
		}
		int q = 0;
		for (size_t r = 0; r < total_rows; r += chunk_rows, q ^= 1) {
			const size_t nr = std::min(chunk_rows, total_rows - r);
			CK(cudaEventSynchronize(done[q]));
			{
				const size_t bytes = nr * row_bytes, nt = 8, per = (bytes + nt - 1) / nt;
#pragma omp parallel for schedule(static) num_threads(8)
				for (long long t = 0; t < (long long)nt; ++t) {
					const size_t b0 = (size_t)t * per, b1 = std::min(bytes, b0 + per);
					if (b1 > b0) memcpy((char*)stage[q] + b0, src + r * row_bytes + b0, b1 - b0);
				}
			}
			CK(cudaMemcpy2DAsync((char*)p + r * (size_t)pitch * ib, (size_t)pitch * ib, stage[q], row_bytes, row_bytes, nr,
			                     cudaMemcpyHostToDevice, stream));
			CK(cudaEventRecord(done[q], stream));
		}
		CK(cudaStreamSynchronize(stream));
		for (int k = 0; k < 2; ++k) { cudaFreeHost(stage[k]); cudaEventDestroy(done[k]); }
	}
	have_compressed = true;
	have_dense = false;
	tmu.lap("index allocation + copy issued");
	return 0;
}

int Engine::set_signal(const float* sv, const float* si, unsigned len, unsigned period)
{
	if (finalized) return fail("engine already finalized");
	if (!sv || !si || len == 0) return fail("set_signal: empty signal");
	h_sig[0].assign(sv, sv + len);
	h_sig[1].assign(si, si + len);
	sig_len = len;
	sig_period = period;
	return 0;
}

int Engine::add_excitation(int is_curr, unsigned count, const unsigned* idx3, const unsigned* dir, const float* amp,
                           const unsigned* delay)
{
	if (finalized) return fail("engine already finalized");
	ExcHost& E = h_exc[is_curr ? 1 : 0];
	for (unsigned n = 0; n < count; ++n) {
		for (int a = 0; a < 3; ++a) {
			if (idx3[(size_t)a * count + n] >= gn[a]) return fail("add_excitation: position outside the mesh");
			E.idx[a].push_back(idx3[(size_t)a * count + n]);
		}
		if (dir[n] > 2) return fail("add_excitation: bad direction");
		E.dir.push_back(dir[n]);
		E.amp.push_back(amp[n]);
		E.delay.push_back(delay[n]);
	}
	return 0;
}

int Engine::add_upml(const unsigned start[3], const unsigned n[3], const float* const c[6])
{
	if (finalized) return fail("engine already finalized");
	if (h_upml.size() >= OEMS_MAX_PML_BOXES) return fail("add_upml: too many boxes");
	UpmlBoxHost B;
	for (int a = 0; a < 3; ++a) {
		if (n[a] == 0 || start[a] + n[a] > gn[a]) return fail("add_upml: box outside the mesh");
		B.start[a] = start[a];
		B.n[a] = n[a];
	}
	const size_t cnt = (size_t)3 * n[0] * n[1] * n[2];
	bool any = false, all = true;
	for (int w = 0; w < 6; ++w) { any |= c[w] != nullptr; all &= c[w] != nullptr; }
	if (any && !all) return fail("add_upml: pass all six coefficient arrays or none");
	if (all)
		for (int w = 0; w < 6; ++w) B.c[w].assign(c[w], c[w] + cnt);
	h_upml.push_back(std::move(B));
	return 0;
}

int Engine::add_mur(int ny, unsigned line, unsigned shift, const unsigned n[2], const float* cP, const float* cPP,
                    unsigned start_ts)
{
	if (finalized) return fail("engine already finalized");
	if (h_mur.size() >= OEMS_MAX_MUR) return fail("add_mur: too many planes");
	if (ny < 0 || ny > 2) return fail("add_mur: bad direction");
	const int nyP = (ny + 1) % 3, nyPP = (ny + 2) % 3;
	if (n[0] != gn[nyP] || n[1] != gn[nyPP] || line >= gn[ny] || shift >= gn[ny]) return fail("add_mur: plane does not match the mesh");
	MurHost M;
	M.ny = ny; M.line = line; M.shift = shift; M.n[0] = n[0]; M.n[1] = n[1]; M.start_ts = start_ts;
	M.cP.assign(cP, cP + (size_t)n[0] * n[1]);
	M.cPP.assign(cPP, cPP + (size_t)n[0] * n[1]);
	h_mur.push_back(std::move(M));
	return 0;
}

int Engine::add_lorentz(unsigned count, const unsigned* pos3, const float* const c[6])
{
	if (finalized) return fail("engine already finalized");
	LorHost L;
	L.count = count;
	L.pos.assign(pos3, pos3 + (size_t)3 * count);
	for (unsigned i = 0; i < count; ++i)
		for (int a = 0; a < 3; ++a)
			if (pos3[(size_t)a * count + i] >= gn[a]) return fail("add_lorentz: position outside the mesh");
	if ((c[0] == nullptr) != (c[1] == nullptr) || (c[3] == nullptr) != (c[4] == nullptr)) return fail("add_lorentz: int/ext coefficients come in pairs");
	if ((c[2] && !c[0]) || (c[5] && !c[3])) return fail("add_lorentz: Lorentz pole without its Drude ADE");
	for (int w = 0; w < 6; ++w)
		if (c[w]) L.c[w].assign(c[w], c[w] + (size_t)3 * count);
	h_lor.push_back(std::move(L));
	return 0;
}

int Engine::add_rlc(unsigned count, const int* dir, const unsigned* pos3, const float* const c[9])
{
	if (finalized) return fail("engine already finalized");
	RlcHost R;
	R.count = count;
	R.dir.assign(dir, dir + count);
	R.pos.assign(pos3, pos3 + (size_t)3 * count);
	for (unsigned i = 0; i < count; ++i) {
		if (dir[i] < 0 || dir[i] > 2) return fail("add_rlc: bad direction");
		for (int a = 0; a < 3; ++a)
			if (pos3[(size_t)a * count + i] >= gn[a]) return fail("add_rlc: position outside the mesh");
	}
	for (int w = 0; w < 9; ++w) {
		if (!c[w]) return fail("add_rlc: null coefficient array");
		R.c[w].assign(c[w], c[w] + count);
	}
	h_rlc.push_back(std::move(R));
	return 0;
}

int Engine::add_steadystate(unsigned period_ts, unsigned count, const unsigned* pos3, const unsigned* dir)
{
	if (finalized) return fail("engine already finalized");
	if (period_ts == 0 || count == 0) return fail("add_steadystate: empty");
	for (unsigned n = 0; n < count; ++n) {
		if (dir[n] > 2) return fail("add_steadystate: bad direction");
		for (int a = 0; a < 3; ++a)
			if (pos3[(size_t)a * count + n] >= gn[a]) return fail("add_steadystate: position outside the mesh");
	}
	ss_period = period_ts;
	// a z-slab engine records the probes on the planes it owns (list order kept) and sums the energy of its planes;
	// the host puts the slabs together (steadystate_raw / steadystate_eval)
	ss_pos.clear(); ss_dir.clear();
	std::vector<unsigned> keep;
	for (unsigned n = 0; n < count; ++n)
		if (owned(pos3[(size_t)2 * count + n])) keep.push_back(n);
	for (int a = 0; a < 3; ++a)
		for (unsigned n : keep) ss_pos.push_back(pos3[(size_t)a * count + n]);
	for (unsigned n : keep) ss_dir.push_back(dir[n]);
	return 0;
}

// Engine_Ext_SteadyState::Apply2Voltages (engine_ext_steadystate.cpp:62-106) evaluated on the host
// from the device snapshot of the last completed period
int Engine::steadystate_raw(unsigned info[2], double en[4], double* snap, unsigned cap, unsigned* count)
{
	if (!ss_on) return fail("steadystate: no steady-state detection set up");
	CK(cudaSetDevice(device));
	const unsigned p = ss_period, cnt = pSs.count;
	if (count) *count = cnt;
	CK(cudaMemcpyAsync(info, pSs.info, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
	CK(cudaMemcpyAsync(en, pSs.energy, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
	if (snap && cnt) {
		if ((size_t)cap < (size_t)2 * p * cnt) return fail("steadystate_raw: buffer too small");
		CK(cudaMemcpyAsync(snap, pSs.snap, (size_t)2 * p * cnt * sizeof(double), cudaMemcpyDeviceToHost, stream));
	}
	CK(cudaStreamSynchronize(stream));
	return 0;
}

// the criterion of Engine_Ext_SteadyState::Apply2Voltages from the raw data; z-slab drivers add up the slabs'
// energies and concatenate their records ([2*period][count], count fastest) before calling it
int Engine::steadystate_eval(unsigned p, unsigned cnt, const unsigned info[2], const double en[4], const double* snap, double* last_diff)
{
	double diff = 1.0;
	if (info[0] > 0) {
		bool no_valid = true;
		diff = 0;
		if (en[2] > 0) { diff = std::fabs(en[3] - en[2]) / en[2]; no_valid = false; }
		const unsigned rel_pos = info[1] % (2 * p);
		unsigned old_pos = 0, new_pos = p;
		if (rel_pos <= p) { new_pos = 0; old_pos = p; }
		std::vector<double> curr_pow(cnt, 0.0), diff_pow(cnt, 0.0);
		double max_pow = 0;
		for (unsigned n = 0; n < cnt; ++n) {
			for (unsigned nt = 0; nt < p; ++nt) {
				const double a = snap[(size_t)(nt + new_pos) * cnt + n], b = snap[(size_t)(nt + old_pos) * cnt + n];
				curr_pow[n] += a * a;
				diff_pow[n] += (b - a) * (b - a);
			}
			max_pow = std::max(max_pow, curr_pow[n]);
		}
		for (unsigned n = 0; n < cnt; ++n)
			if (curr_pow[n] > max_pow * 1e-2) { diff = std::max(diff, diff_pow[n] / curr_pow[n]); no_valid = false; }
		if (no_valid || diff > 1) diff = 1;
	}
	if (last_diff) *last_diff = diff;
	return 0;
}

// Engine_Ext_SteadyState::Apply2Voltages (engine_ext_steadystate.cpp:62-106) evaluated on the host
// from the device snapshot of the last completed period
int Engine::steadystate_check(double* last_diff, unsigned* n_checks)
{
	if (!ss_on) return fail("steadystate_check: no steady-state detection set up");
	if (slab_set && (zb > 0 || ze < gn[2]))
		return fail("steadystate_check: on z-slab engines gather oems_cuda_steadystate_raw of all slabs and call oems_cuda_steadystate_eval");
	const unsigned p = ss_period, cnt = pSs.count;
	unsigned info[2];
	double en[4];
	std::vector<double> snap((size_t)2 * p * cnt);
	if (steadystate_raw(info, en, snap.data(), (unsigned)snap.size(), nullptr)) return 1;
	if (n_checks) *n_checks = info[0];
	return steadystate_eval(p, cnt, info, en, snap.data(), last_diff);
}

// ------------------------------------------------------------------------------ compression
// Re-keys the operator per cell: the 12 stencil coefficients plus, inside UPML
// boxes, the 18 auxiliary coefficients form one 128-byte tuple; equal tuples (memcmp, like
// SSE_coeff operator_sse_compressed.cpp:197-200) share an entry.
#include "entry_set.h"

int Engine::compress_dense(std::vector<uint32_t>& index32)
{
	const unsigned nx = gn[0], ny = gn[1];
	index32.assign((size_t)nx * ny * nzl, 0);
	const int nthreads = std::max(1, std::min(omp_get_max_threads(), nzl));
	std::vector<EntrySet> sets(nthreads);
	std::vector<int> owner(nzl, 0); // the thread (= thread-local set) that keyed plane kl: the remap below must not rely on the runtime handing out the same planes again
	for (auto& B : h_upml)
		if (B.c[0].empty()) return fail("dense operator needs dense UPML coefficient arrays");
#pragma omp parallel num_threads(nthreads)
	{
		const int t = omp_get_thread_num() % nthreads;
		EntrySet& S = sets[t];
#pragma omp for schedule(static)
		for (int kl = 0; kl < nzl; ++kl) {
			owner[kl] = t;
			const unsigned gz = (unsigned)(z0 + kl);
			for (unsigned j = 0; j < ny; ++j)
				for (unsigned i = 0; i < nx; ++i) {
					oems_coeff_entry e;
					memset(&e, 0, sizeof(e));
					for (int n = 0; n < 3; ++n) {
						const size_t o = (((size_t)n * nx + i) * ny + j) * nzl + kl;
						e.vv[n] = h_dense[0][o]; e.vi[n] = h_dense[1][o];
						e.ii[n] = h_dense[2][o]; e.iv[n] = h_dense[3][o];
					}
					for (const auto& B : h_upml) {
						const unsigned li = i - B.start[0], lj = j - B.start[1], lk = gz - B.start[2];
						if (li < B.n[0] && lj < B.n[1] && lk < B.n[2]) {
							e.pml = 1.0f;
							for (int n = 0; n < 3; ++n) {
								const size_t o = (((size_t)n * B.n[0] + li) * B.n[1] + lj) * B.n[2] + lk;
								e.pml_vv[n] = B.c[0][o]; e.pml_vvfn[n] = B.c[1][o]; e.pml_vvfo[n] = B.c[2][o];
								e.pml_ii[n] = B.c[3][o]; e.pml_iifn[n] = B.c[4][o]; e.pml_iifo[n] = B.c[5][o];
							}
							break;
						}
					}
					// thread-local id, tagged with the thread in the upper bits for the merge
					index32[((size_t)kl * ny + j) * nx + i] = S.insert(e);
				}
		}
	}
	// merge the thread-local sets
	EntrySet G;
	std::vector<std::vector<uint32_t>> remap(nthreads);
	for (int t = 0; t < nthreads; ++t) {
		remap[t].resize(sets[t].items.size());
		for (size_t u = 0; u < sets[t].items.size(); ++u) remap[t][u] = G.insert(sets[t].items[u]);
	}
#pragma omp parallel for schedule(static)
	for (int kl = 0; kl < nzl; ++kl) {
		uint32_t* row = index32.data() + (size_t)kl * ny * nx;
		const std::vector<uint32_t>& R = remap[owner[kl]];
		for (size_t q = 0; q < (size_t)ny * nx; ++q) row[q] = R[row[q]];
	}
	h_table.swap(G.items);
	n_unique = (unsigned)h_table.size();
	for (int a = 0; a < 4; ++a) std::vector<float>().swap(h_dense[a]);
	return 0;
}

int Engine::build_tables_and_index(const std::vector<uint32_t>& index32)
{
	const unsigned U = n_unique;
	// device tables: U real entries + the all-zero entry U used by padding cells
	std::vector<float4> t[10];
	for (int w = 0; w < 10; ++w) t[w].assign(U + 1, make_float4(0, 0, 0, 0));
	for (unsigned u = 0; u < U; ++u) {
		const oems_coeff_entry& e = h_table[u];
		const float flag = e.pml != 0.0f ? 1.0f : 0.0f;
		t[0][u] = make_float4(e.vv[0], e.vv[1], e.vv[2], flag);
		t[1][u] = make_float4(e.vi[0], e.vi[1], e.vi[2], 0);
		t[2][u] = make_float4(e.pml_vv[0], e.pml_vv[1], e.pml_vv[2], 0);
		t[3][u] = make_float4(e.pml_vvfn[0], e.pml_vvfn[1], e.pml_vvfn[2], 0);
		t[4][u] = make_float4(e.pml_vvfo[0], e.pml_vvfo[1], e.pml_vvfo[2], 0);
		t[5][u] = make_float4(e.ii[0], e.ii[1], e.ii[2], flag);
		t[6][u] = make_float4(e.iv[0], e.iv[1], e.iv[2], 0);
		t[7][u] = make_float4(e.pml_ii[0], e.pml_ii[1], e.pml_ii[2], 0);
		t[8][u] = make_float4(e.pml_iifn[0], e.pml_iifn[1], e.pml_iifn[2], 0);
		t[9][u] = make_float4(e.pml_iifo[0], e.pml_iifo[1], e.pml_iifo[2], 0);
	}
	for (int w = 0; w < 10; ++w) {
		d_tab[w] = upload(t[w]);
		if (!d_tab[w]) return fail("out of device memory (tables)");
	}
	if (have_compressed) return 0; // index already in HBM
	index_bytes = (U + 1 <= 65536) ? 2 : 4;
	const size_t cells = (size_t)plane * nzl;
	const unsigned nx = gn[0], ny = gn[1];
	if (index_bytes == 2) {
		std::vector<uint16_t> h(cells, (uint16_t)U);
#pragma omp parallel for schedule(static)
		for (long long r = 0; r < (long long)ny * nzl; ++r)
			for (unsigned i = 0; i < nx; ++i) h[(size_t)r * pitch + i] = (uint16_t)index32[(size_t)r * nx + i];
		d_idx = upload(h);
		CK(cudaStreamSynchronize(stream));
	} else {
		std::vector<uint32_t> h(cells, U);
#pragma omp parallel for schedule(static)
		for (long long r = 0; r < (long long)ny * nzl; ++r)
			for (unsigned i = 0; i < nx; ++i) h[(size_t)r * pitch + i] = index32[(size_t)r * nx + i];
		d_idx = upload(h);
		CK(cudaStreamSynchronize(stream));
	}
	if (!d_idx) return fail("out of device memory (index)");
	return 0;
}

// ------------------------------------------------------------------------------ extensions
int Engine::build_pml()
{
	flux_floats = 0;
	pml_cells = 0;
	int nb = 0;
	edge_possible = false;
	for (auto& B : h_upml) {
		// part of the box on the planes this GPU updates (owned planes)
		const unsigned bz0 = std::max(B.start[2], zb), bz1 = std::min(B.start[2] + B.n[2], ze);
		B.ln[2] = 0;
		if (bz1 <= bz0) continue;
		B.ls[0] = (int)B.start[0]; B.ls[1] = (int)B.start[1]; B.ls[2] = (int)bz0 - z0;
		B.ln[0] = (int)B.n[0]; B.ln[1] = (int)B.n[1]; B.ln[2] = (int)(bz1 - bz0);
		B.gz0 = bz0;
		B.flux_off = flux_floats;
		const long long cs = (long long)B.ln[0] * B.ln[1] * B.ln[2];
		flux_floats += 3 * cs;
		pml_cells += (uint64_t)cs;
		PmlBox pb;
		for (int a = 0; a < 3; ++a) { pb.s[a] = B.ls[a]; pb.n[a] = B.ln[a]; }
		pb.off = B.flux_off;
		pE.box[nb] = pb;
		pH.box[nb] = pb;
		++nb;
		// cells of the box the H stencil never visits (last line of a direction): listed only when
		// somebody writes a current there (build_edge_list)
		if (B.ls[0] + B.ln[0] == (int)gn[0] || B.ls[1] + B.ln[1] == (int)gn[1] || (int)bz0 + B.ln[2] == (int)gn[2]) edge_possible = true;
	}
	pE.nboxes = pH.nboxes = nb;
	has_pml = nb > 0;
	pml_disjoint = true;
	for (int a = 0; a < nb; ++a)
		for (int b = a + 1; b < nb; ++b) {
			bool overlap = true;
			for (int d = 0; d < 3; ++d)
				overlap &= pE.box[a].s[d] < pE.box[b].s[d] + pE.box[b].n[d] && pE.box[b].s[d] < pE.box[a].s[d] + pE.box[a].n[d];
			if (overlap) pml_disjoint = false;
		}
	if (has_pml) {
		d_flux_v = dalloc<float>((size_t)flux_floats);
		d_flux_i = dalloc<float>((size_t)flux_floats);
		if (!d_flux_v || !d_flux_i) return fail("out of device memory (UPML flux)");
	}
	pEdge.count = 0;
	return 0;
}

// UPML cells the H stencil never visits (k_upml_untouched_H): built the first time a current is
// written there -- at 1024^3 the list has 6.3 M entries and costs 0.1 s that no ordinary run needs
int Engine::build_edge_list()
{
	std::vector<long long> e_cell, e_fo, e_cs;
	for (auto& B : h_upml) {
		if (B.ln[2] <= 0) continue;
		const unsigned bz0 = B.gz0;
		const long long cs = (long long)B.ln[0] * B.ln[1] * B.ln[2];
		auto add_edge = [&](int li, int lj, int lk) {
			e_cell.push_back(cell_off(B.ls[0] + li, B.ls[1] + lj, bz0 + lk));
			e_fo.push_back(B.flux_off + ((long long)lk * B.ln[1] + lj) * B.ln[0] + li);
			e_cs.push_back(cs);
		};
		const int lx = (int)gn[0] - 1 - B.ls[0], ly = (int)gn[1] - 1 - B.ls[1], lz = (int)gn[2] - 1 - (int)bz0;
		const bool hx = lx >= 0 && lx < B.ln[0], hy = ly >= 0 && ly < B.ln[1], hz = lz >= 0 && lz < B.ln[2];
		if (hx)
			for (int lk = 0; lk < B.ln[2]; ++lk)
				for (int lj = 0; lj < B.ln[1]; ++lj) add_edge(lx, lj, lk);
		if (hy)
			for (int lk = 0; lk < B.ln[2]; ++lk)
				for (int li = 0; li < B.ln[0]; ++li)
					if (!(hx && li == lx)) add_edge(li, ly, lk);
		if (hz)
			for (int lj = 0; lj < B.ln[1]; ++lj)
				for (int li = 0; li < B.ln[0]; ++li)
					if (!(hx && li == lx) && !(hy && lj == ly)) add_edge(li, lj, lz);
	}
	pEdge.count = (long long)e_cell.size();
	if (pEdge.count) {
		pEdge.cell = upload(e_cell);
		pEdge.fluxoff = upload(e_fo);
		pEdge.fluxcs = upload(e_cs);
		pEdge.X = d_I;
		pEdge.idx = d_idx;
		pEdge.tP0 = d_tab[7]; pEdge.tP1 = d_tab[8]; pEdge.tP2 = d_tab[9];
		pEdge.flux = d_flux_i;
		pEdge.comp = comp;
		if (!pEdge.cell || !pEdge.fluxoff || !pEdge.fluxcs) return fail("out of device memory (UPML edge list)");
	}
	return 0;
}

// ------------------------------------------------------------------------------ TFSF
int Engine::set_tfsf(const unsigned start[3], const unsigned stop[3], const int active[6], const unsigned* const* vdelay,
                     const float* const* vdd, const float* const* vamp, const unsigned* const* cdelay, const float* const* cdd,
                     const float* const* camp)
{
	if (finalized) return fail("engine already finalized");
	TfsfHost& T = h_tfsf;
	for (int n = 0; n < 3; ++n) {
		if (start[n] > stop[n] || stop[n] >= gn[n]) return fail("set_tfsf: box outside the mesh");
		T.start[n] = start[n]; T.stop[n] = stop[n];
		T.active[n][0] = active[2 * n] != 0; T.active[n][1] = active[2 * n + 1] != 0;
		if (T.active[n][0] && start[n] == 0) return fail("set_tfsf: an active low face needs a line below it");
	}
	for (int n = 0; n < 3; ++n) {
		const size_t numP = (size_t)(stop[(n + 1) % 3] - start[(n + 1) % 3] + 1) * (stop[(n + 2) % 3] - start[(n + 2) % 3] + 1);
		for (int l = 0; l < 2; ++l)
			for (int c = 0; c < 2; ++c) {
				const int q = (n * 2 + l) * 2 + c;
				if (!T.active[n][l]) continue;
				if (!vdelay[q] || !vdd[q] || !vamp[q] || !cdelay[q] || !cdd[q] || !camp[q]) return fail("set_tfsf: missing table of an active face");
				T.delay[0][q].assign(vdelay[q], vdelay[q] + numP); T.dd[0][q].assign(vdd[q], vdd[q] + numP); T.amp[0][q].assign(vamp[q], vamp[q] + numP);
				T.delay[1][q].assign(cdelay[q], cdelay[q] + numP); T.dd[1][q].assign(cdd[q], cdd[q] + numP); T.amp[1][q].assign(camp[q], camp[q] + numP);
			}
	}
	T.on = true;
	return 0;
}

int Engine::build_tfsf()
{
	memset(pTfsf, 0, sizeof(pTfsf));
	const TfsfHost& T = h_tfsf;
	if (!T.on) return 0;
	if (sig_len == 0) return fail("TFSF without a signal (set_signal)");
	for (int w = 0; w < 2; ++w) {
		// the reference's loop nest: n, lower then upper face, i over nP, j over nPP, component nP then nPP
		std::map<long long, std::vector<std::pair<int, unsigned>>> groups; // target -> (table, point) in order
		std::vector<long long> order;
		for (int n = 0; n < 3; ++n) {
			const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
			const unsigned nl0 = T.stop[nP] - T.start[nP] + 1, nl1 = T.stop[nPP] - T.start[nPP] + 1;
			for (int l = 0; l < 2; ++l) {
				if (!T.active[n][l]) continue;
				unsigned u = 0;
				for (unsigned i = 0; i < nl0; ++i)
					for (unsigned j = 0; j < nl1; ++j, ++u) {
						unsigned pos[3];
						pos[nP] = T.start[nP] + i; pos[nPP] = T.start[nPP] + j;
						pos[n] = l ? T.stop[n] : (w ? T.start[n] - 1 : T.start[n]);
						if (!owned(pos[2])) continue;
						for (int c = 0; c < 2; ++c) {
							const long long t = (long long)(c ? nPP : nP) * comp + cell_off(pos[0], pos[1], pos[2]);
							auto it = groups.find(t);
							if (it == groups.end()) { groups[t] = {{(n * 2 + l) * 2 + c, u}}; order.push_back(t); }
							else it->second.push_back({(n * 2 + l) * 2 + c, u});
						}
					}
			}
		}
		if (order.empty()) continue;
		std::vector<long long> tgt;
		std::vector<unsigned> gstart, delay;
		std::vector<float> dd, amp;
		for (long long t : order) {
			tgt.push_back(t);
			gstart.push_back((unsigned)amp.size());
			for (auto& e : groups[t]) { delay.push_back(T.delay[w][e.first][e.second]); dd.push_back(T.dd[w][e.first][e.second]); amp.push_back(T.amp[w][e.first][e.second]); }
		}
		gstart.push_back((unsigned)amp.size());
		TfsfParams& P = pTfsf[w];
		P.X = w ? d_I : d_V;
		P.tgt = upload(tgt); P.gstart = upload(gstart); P.delay = upload(delay); P.dd = upload(dd); P.amp = upload(amp);
		P.sig = d_sig[w ? 0 : 1]; // "get the current signal since an H-field is added" and vice versa
		P.numTS = d_numTS;
		P.groups = (unsigned)tgt.size(); P.length = sig_len; P.period = sig_period;
		if (!P.tgt || !P.gstart || !P.delay || !P.dd || !P.amp) return fail("out of device memory (TFSF)");
	}
	return 0;
}

// ------------------------------------------------------------------------------ absorbing sheets
int Engine::add_absorbing_sheet(int ny, const unsigned x0[3], const unsigned x1[3], int positive, int type, const float* K1P,
                                const float* K1PP, const float* K2P, const float* K2PP)
{
	if (finalized) return fail("engine already finalized");
	if (h_sheet.size() >= 8) return fail("add_absorbing_sheet: too many sheets");
	if (ny < 0 || ny > 2 || (type != 1 && type != 2) || !K1P || !K1PP) return fail("add_absorbing_sheet: bad arguments");
	if (type == 2 && (!K2P || !K2PP)) return fail("add_absorbing_sheet: super-absorption needs the K2 arrays");
	SheetHost S;
	S.ny = ny; S.type = type; S.positive = positive != 0;
	for (int a = 0; a < 3; ++a) {
		if (x0[a] > x1[a] || x1[a] >= gn[a]) return fail("add_absorbing_sheet: sheet outside the mesh");
		S.x0[a] = x0[a]; S.x1[a] = x1[a];
	}
	if (x0[ny] != x1[ny]) return fail("add_absorbing_sheet: not a sheet normal to ny");
	// shifted lines of the engine extension ctor (engine_ext_absorbing_bc.cpp:64-71) must exist
	const long long line = x0[ny], lo = S.positive ? line : line - (type == 2 ? 2 : 1), hi = S.positive ? line + 1 : line;
	if (lo < 0 || hi >= (long long)gn[ny]) return fail("add_absorbing_sheet: the shifted line is outside the mesh");
	const int nP = (ny + 1) % 3, nPP = (ny + 2) % 3;
	const size_t n = (size_t)(x1[nP] - x0[nP] + 1) * (x1[nPP] - x0[nPP] + 1);
	S.K1P.assign(K1P, K1P + n); S.K1PP.assign(K1PP, K1PP + n);
	if (type == 2) { S.K2P.assign(K2P, K2P + n); S.K2PP.assign(K2PP, K2PP + n); }
	h_sheet.push_back(std::move(S));
	return 0;
}

int Engine::build_sheets()
{
	sheet_dev.clear();
	if (h_sheet.empty()) return 0;
	for (const SheetHost& S : h_sheet) {
		const int ny = S.ny, nP = (ny + 1) % 3, nPP = (ny + 2) % 3;
		const unsigned nl0 = S.x1[nP] - S.x0[nP] + 1, nl1 = S.x1[nPP] - S.x0[nPP] + 1;
		const unsigned line = S.x0[ny];
		const unsigned shift_V = line + (S.positive ? 1 : -1);
		const unsigned pos_I = line + (S.positive ? 0 : -1), shift_I = line + (S.positive ? 1 : -2);
		// z-slab engines (template: the extension simply lives on the MPI rank that holds the cells): sheets normal to
		// x or y are cut at the slab's owned planes; a sheet normal to z belongs to the slab that owns its line, and
		// the lines it reads next to it must be owned by the same slab
		bool mine = true;
		if (slab_set && ny == 2) {
			mine = owned(line);
			const bool sa = S.type == 2 && nl0 > 1 && nl1 > 1;
			if (mine && (!owned(shift_V) || (sa && (!owned(pos_I) || !owned(shift_I)))))
				return fail("absorbing sheet normal to z lies on a z-slab boundary: move the split by a few planes");
			if (!mine && (owned(shift_V) || (sa && (owned(pos_I) || owned(shift_I)))))
				return fail("absorbing sheet normal to z lies on a z-slab boundary: move the split by a few planes");
		}
		SheetDev D;
		memset(&D, 0, sizeof(D));
		auto off = [&](int comp_n, unsigned l, unsigned a, unsigned b) {
			unsigned pos[3];
			pos[ny] = l; pos[nP] = S.x0[nP] + a; pos[nPP] = S.x0[nPP] + b;
			return (long long)comp_n * comp + cell_off(pos[0], pos[1], pos[2]);
		};
		auto here = [&](unsigned a, unsigned b) {
			if (!slab_set) return true;
			if (ny == 2) return mine;
			const unsigned z = nP == 2 ? S.x0[nP] + a : S.x0[nPP] + b;
			return owned(z);
		};
		// voltage list: all sheet points, component nyP then nyPP of a point (order is irrelevant: distinct cells)
		std::vector<long long> o, os;
		std::vector<float> k1, k2;
		for (unsigned a = 0; a < nl0; ++a)
			for (unsigned b = 0; b < nl1; ++b) {
				if (!here(a, b)) continue;
				const size_t q = (size_t)a * nl1 + b;
				o.push_back(off(nP, line, a, b)); os.push_back(off(nP, shift_V, a, b)); k1.push_back(S.K1P[q]);
				o.push_back(off(nPP, line, a, b)); os.push_back(off(nPP, shift_V, a, b)); k1.push_back(S.K1PP[q]);
			}
		D.v.count = (long long)o.size();
		D.v.o = upload(o); D.v.os = upload(os); D.v.K1 = upload(k1);
		D.v.store = dalloc<float>(o.size());
		D.v.X = d_V;
		if (S.type == 2 && nl0 > 1 && nl1 > 1) {
			o.clear(); os.clear(); k1.clear();
			for (unsigned a = 0; a + 1 < nl0; ++a)
				for (unsigned b = 0; b + 1 < nl1; ++b) {
					if (!here(a, b)) continue;
					const size_t q = (size_t)a * nl1 + b;
					o.push_back(off(nP, pos_I, a, b)); os.push_back(off(nP, shift_I, a, b)); k1.push_back(S.K1P[q]); k2.push_back(S.K2P[q]);
					o.push_back(off(nPP, pos_I, a, b)); os.push_back(off(nPP, shift_I, a, b)); k1.push_back(S.K1PP[q]); k2.push_back(S.K2PP[q]);
				}
			D.i.count = (long long)o.size();
			D.i.o = upload(o); D.i.os = upload(os); D.i.K1 = upload(k1); D.i.K2 = upload(k2);
			D.i.store = dalloc<float>(o.size());
			D.i.X = d_I;
		}
		sheet_dev.push_back(D);
	}
	return 0;
}

// no voltage excitation entry writes a component that a Mur plane writes (the two tangential components on the plane's
// boundary line): Apply2Voltages of the two extensions then commute and share a launch (k_mur_apply_excite)
bool Engine::mur_exc_disjoint() const
{
	const ExcHost& E = h_exc[0];
	for (size_t n = 0; n < E.dir.size(); ++n) {
		const unsigned pos[3] = {E.idx[0][n], E.idx[1][n], E.idx[2][n]};
		for (const MurHost& M : h_mur)
			if (pos[M.ny] == M.line && (int)E.dir[n] != M.ny) return false;
	}
	return true;
}

int Engine::build_mur()
{
	memset(&pMur, 0, sizeof(pMur));
	if (h_mur.empty()) return 0;
	long long total = 0;
	std::vector<float> cP, cPP;
	for (size_t m = 0; m < h_mur.size(); ++m) {
		const MurHost& M = h_mur[m];
		MurPlane& P = pMur.pl[m];
		P.ny = M.ny; P.nyP = (M.ny + 1) % 3; P.nyPP = (M.ny + 2) % 3;
		P.line = (int)M.line; P.shift = (int)M.shift;
		P.n0 = (int)M.n[0]; P.n1 = (int)M.n[1];
		P.start_ts = M.start_ts;
		P.eoff = total;
		total += (long long)M.n[0] * M.n[1];
		cP.insert(cP.end(), M.cP.begin(), M.cP.end());
		cPP.insert(cPP.end(), M.cPP.begin(), M.cPP.end());
	}
	// write conflicts on shared edges: Apply2Voltages runs the planes in reverse insertion order
	// (engine.cpp:87-91,239-244), so the plane inserted FIRST writes last and wins -- once it is active
	// (IsActive(), engine_ext_mur_abc.h:56).  A component c of a cell is written by at most two planes
	// (normals != c).  Per entry and component: the start timestep of the earlier-inserted plane that
	// overrides this write (0xffffffff: nobody does); resolved at run time against numTS, so planes with
	// different start delays (a source on one Mur face) follow the reference step by step.
	std::vector<unsigned> ovr((size_t)2 * total, 0xffffffffu);
	for (size_t m = 0; m < h_mur.size(); ++m) {
		const MurPlane& P = pMur.pl[m];
		for (int a = 0; a < P.n0; ++a)
			for (int b = 0; b < P.n1; ++b) {
				int pos[3];
				pos[P.ny] = P.line; pos[P.nyP] = a; pos[P.nyPP] = b;
				const int comps[2] = {P.nyP, P.nyPP};
				for (int q = 0; q < 2; ++q) {
					const int other = 3 - P.ny - comps[q]; // normal of the other plane that writes comps[q]
					for (size_t m2 = 0; m2 < m; ++m2)
						if (pMur.pl[m2].ny == other && pos[other] == pMur.pl[m2].line)
							ovr[(size_t)q * total + (size_t)P.eoff + (size_t)a * P.n1 + b] = pMur.pl[m2].start_ts;
				}
			}
	}
	pMur.V = d_V;
	pMur.cP = upload(cP); pMur.cPP = upload(cPP);
	pMur.vP = dalloc<float>((size_t)total); pMur.vPP = dalloc<float>((size_t)total);
	pMur.ovr_start = upload(ovr);
	pMur.numTS = d_numTS;
	pMur.nplanes = (int)h_mur.size();
	pMur.total = total;
	pMur.pitch = pitch; pMur.plane = plane; pMur.comp = comp;
	pMur.z0 = z0; pMur.zown0 = (int)zb; pMur.zown1 = (int)ze;
	if (!pMur.cP || !pMur.cPP || !pMur.vP || !pMur.vPP || !pMur.ovr_start) return fail("out of device memory (Mur)");
	return 0;
}

int Engine::build_exc()
{
	for (int w = 0; w < 2; ++w) {
		memset(&pExc[w], 0, sizeof(ExcParams));
		const ExcHost& E = h_exc[w];
		const size_t cnt = E.dir.size();
		if (cnt == 0) continue;
		if (sig_len == 0) return fail("excitation without a signal (set_signal)");
		// group by target, keep list order inside a group (sequential += of the CPU loop)
		std::map<long long, std::vector<unsigned>> groups;
		std::vector<long long> order;
		for (unsigned n = 0; n < cnt; ++n) {
			if (!owned(E.idx[2][n])) continue;
			const long long t = (long long)E.dir[n] * comp + cell_off(E.idx[0][n], E.idx[1][n], E.idx[2][n]);
			auto it = groups.find(t);
			if (it == groups.end()) { groups[t] = {n}; order.push_back(t); }
			else it->second.push_back(n);
		}
		if (order.empty()) continue;
		std::vector<long long> tgt;
		std::vector<unsigned> gstart, delay;
		std::vector<float> amp;
		for (long long t : order) {
			tgt.push_back(t);
			gstart.push_back((unsigned)amp.size());
			for (unsigned n : groups[t]) { amp.push_back(E.amp[n]); delay.push_back(E.delay[n]); }
		}
		gstart.push_back((unsigned)amp.size());
		ExcParams& P = pExc[w];
		P.X = w ? d_I : d_V;
		P.tgt = upload(tgt); P.gstart = upload(gstart); P.amp = upload(amp); P.delay = upload(delay);
		P.sig = d_sig[w];
		P.numTS = d_numTS;
		P.groups = (unsigned)tgt.size();
		P.length = sig_len;
		P.period = sig_period;
		if (!P.tgt || !P.gstart || !P.amp || !P.delay) return fail("out of device memory (excitation)");
	}
	return 0;
}

// Can the one-pass kernel apply the ADE of the Lorentz/Drude cells itself?  It computes E_new = stencil - ADE and
// H_new = stencil - ADE per cell (engine_ext_lorentzmaterial.cpp:79-168: the pre hooks only read timestep-n values and
// run as list kernels before it).  That reproduces the reference's hook order as long as no OTHER hook reads or
// changes a dispersive cell between the stencil and Apply2Voltages / Apply2Current:
//   * at most two orders (poles); in every row the cells of an order are contiguous in x (one segment per row);
//   * no dispersive cell inside a UPML box (its post-voltage hook precedes the ADE subtraction) or in the float4
//     chunks / 16-line window such a box touches in x (plain cells there are updated by the shell / window kernels);
//   * none on a Mur plane's boundary or shifted line (k_mur_post reads the stencil value before the subtraction),
//     no absorbing sheets and no TFSF box together with dispersive material;
//   * no H cell of the fix-up list is dispersive (checked with the list, build_fix_list).
// Everything else runs the two-pass schedule.
bool Engine::lorentz_fusable() const
{
	if (h_lor.size() > LOR_FUSED_MAX) return false;
	if (!h_sheet.empty() || h_tfsf.on) return false;
	for (const LorHost& L : h_lor) {
		const unsigned* px = L.pos.data();
		const unsigned* py = px + L.count;
		const unsigned* pz = py + L.count;
		for (unsigned i = 0; i < L.count; ++i) {
			for (const UpmlBoxHost& B : h_upml) {
				// in x the shell launches also own the plain cells of the float4 chunks a box touches (the 16-line
				// windows of k_xslab_tma are not used next to dispersive cells: build_schedule_fused)
				const unsigned lo = B.start[0] / 4 * 4, hi = (B.start[0] + B.n[0] + 3) / 4 * 4;
				if (px[i] >= lo && px[i] < hi && py[i] - B.start[1] < B.n[1] && pz[i] - B.start[2] < B.n[2]) return false;
			}
			const unsigned pos[3] = {px[i], py[i], pz[i]};
			for (const MurHost& M : h_mur)
				if (pos[M.ny] == M.line || pos[M.ny] == M.shift) return false;
		}
	}
	return true;
}

int Engine::build_lorentz()
{
	lor_dev.clear();
	lor_xmin = 1 << 30; lor_xmax = -1;
	for (const LorHost& L : h_lor) {
		// keep only the cells on owned planes, in storage order (z, y, x): the list kernels then read the fields
		// coalesced, and the one-pass kernel finds a cell's ADE through one {first x, cells, first index} per row
		std::vector<std::pair<long long, unsigned>> kv; // (storage offset, list entry) of the owned cells
		for (unsigned i = 0; i < L.count; ++i)
			if (owned(L.pos[(size_t)2 * L.count + i]))
				kv.emplace_back(cell_off(L.pos[i], L.pos[(size_t)L.count + i], L.pos[(size_t)2 * L.count + i]), i);
		std::sort(kv.begin(), kv.end());
		const unsigned cnt = (unsigned)kv.size();
		std::vector<unsigned> keep(cnt);
		std::vector<long long> cell(cnt);
		for (unsigned q = 0; q < cnt; ++q) { cell[q] = kv[q].first; keep[q] = kv[q].second; }
		std::vector<std::pair<long long, unsigned>>().swap(kv);
		for (unsigned i : keep) { lor_xmin = std::min(lor_xmin, (int)L.pos[i]); lor_xmax = std::max(lor_xmax, (int)L.pos[i]); }
		for (unsigned q = 1; q < cnt; ++q)
			if (cell[q] == cell[q - 1]) return fail("add_lorentz: a cell is listed twice");
		auto pick = [&](const std::vector<float>& src) {
			std::vector<float> d((size_t)3 * cnt);
			for (int n = 0; n < 3; ++n) {
#pragma omp parallel for schedule(static)
				for (long long q = 0; q < (long long)cnt; ++q) d[(size_t)n * cnt + q] = src[(size_t)n * L.count + keep[q]];
			}
			return d;
		};
		LorDev D;
		memset(&D.v, 0, sizeof(LorParams));
		memset(&D.i, 0, sizeof(LorParams));
		D.v_on = !L.c[0].empty() && cnt > 0;
		D.i_on = !L.c[3].empty() && cnt > 0;
		const long long* d_cell = cnt ? upload(cell) : nullptr;
		if (D.v_on) {
			D.v.X = d_V; D.v.cell = d_cell; D.v.count = cnt; D.v.comp = comp;
			D.v.c_int = upload(pick(L.c[0])); D.v.c_ext = upload(pick(L.c[1]));
			D.v.ade = dalloc<float>((size_t)3 * cnt);
			if (!L.c[2].empty()) { D.v.c_lor = upload(pick(L.c[2])); D.v.lor_ade = dalloc<float>((size_t)3 * cnt); }
		}
		if (D.i_on) {
			D.i.X = d_I; D.i.cell = d_cell; D.i.count = cnt; D.i.comp = comp;
			D.i.c_int = upload(pick(L.c[3])); D.i.c_ext = upload(pick(L.c[4]));
			D.i.ade = dalloc<float>((size_t)3 * cnt);
			if (!L.c[5].empty()) { D.i.c_lor = upload(pick(L.c[5])); D.i.lor_ade = dalloc<float>((size_t)3 * cnt); }
		}
		// coefficient compression (ADE coefficients table-compressed): the distinct {int, ext, lor} x 3 tuples
		// of a list -- interior, faces, edges, corners of a homogeneous block -- form a small table, the list kernels read
		// a 16-bit index per cell instead of 24 .. 36 bytes of coefficients
		auto compress = [&](LorParams& P, const std::vector<float>& ci, const std::vector<float>& ce, const std::vector<float>& cl) {
			P.pidx = nullptr; P.ptab = nullptr;
			if (!cnt) return;
			std::map<std::array<uint32_t, 9>, unsigned> seen;
			std::array<uint32_t, 9> last_key{};
			std::vector<unsigned short> idx(cnt);
			std::vector<float> tab;
			for (unsigned q = 0; q < cnt; ++q) {
				std::array<uint32_t, 9> key;
				float f[9];
				for (int n = 0; n < 3; ++n) {
					const size_t i = (size_t)n * L.count + keep[q];
					f[n] = ci[i]; f[3 + n] = ce[i]; f[6 + n] = cl.empty() ? 0.0f : cl[i];
				}
				memcpy(key.data(), f, sizeof(f));
				if (q && key == last_key) { idx[q] = idx[q - 1]; continue; } // runs of equal cells: no lookup
				last_key = key;
				auto it = seen.find(key);
				if (it == seen.end()) {
					if (seen.size() >= 65536) return; // too many: keep the per-cell arrays
					it = seen.emplace(key, (unsigned)seen.size()).first;
					tab.insert(tab.end(), f, f + 9);
				}
				idx[q] = (unsigned short)it->second;
			}
			P.pidx = upload(idx); P.ptab = upload(tab);
			if (!P.pidx || !P.ptab) { P.pidx = nullptr; P.ptab = nullptr; }
		};
		if (D.v_on) compress(D.v, L.c[0], L.c[1], L.c[2]);
		if (D.i_on) compress(D.i, L.c[3], L.c[4], L.c[5]);
		// entries on the slab's top owned plane (its H is done after the neighbour's E plane has arrived)
		D.top_first = cnt;
		{
			const long long top = (long long)((int)ze - 1 - z0) * plane;
			D.top_first = (unsigned)(std::lower_bound(cell.begin(), cell.end(), top) - cell.begin());
		}
		if (lor_fused) {
			std::vector<int2> rows((size_t)nzl * gn[1], make_int2(0, 0));
			for (unsigned q = 0; q < cnt && lor_fused; ++q) {
				const long long r = cell[q] / pitch;
				const int x = (int)(cell[q] % pitch);
				int2& R = rows[(size_t)r];
				const int n = R.x >> 16, x0 = R.x & 0xffff;
				if (n == 0) { R.x = x | (1 << 16); R.y = (int)q; }
				else if (x == x0 + n && n < 32767) R.x = x0 | ((n + 1) << 16);
				else lor_fused = false; // a second segment in this row
			}
			if (gn[0] > 65535) lor_fused = false;
			if (lor_fused) D.d_rows = upload(rows);
			if (lor_fused && !D.d_rows) return fail("out of device memory (Lorentz row table)");
			if (lor_fused) D.h_rows.swap(rows);
		}
		lor_dev.push_back(D);
	}
	if (!h_lor.empty() && !lor_fused) fused_possible = false;
	return 0;
}

int Engine::build_rlc()
{
	rlc_dev.clear();
	for (const RlcHost& R : h_rlc) {
		std::vector<unsigned> keep;
		for (unsigned i = 0; i < R.count; ++i)
			if (owned(R.pos[(size_t)2 * R.count + i])) keep.push_back(i);
		const unsigned cnt = (unsigned)keep.size();
		if (!cnt) continue;
		std::vector<long long> tgt(cnt);
		for (unsigned q = 0; q < cnt; ++q) {
			const unsigned i = keep[q];
			tgt[q] = (long long)R.dir[i] * comp + cell_off(R.pos[i], R.pos[(size_t)R.count + i], R.pos[(size_t)2 * R.count + i]);
		}
		auto pick = [&](const std::vector<float>& src) {
			std::vector<float> d(cnt);
			for (unsigned q = 0; q < cnt; ++q) d[q] = src[keep[q]];
			return d;
		};
		RlcParams P;
		memset(&P, 0, sizeof(P));
		P.V = d_V;
		P.tgt = upload(tgt);
		P.ilv = upload(pick(R.c[0])); P.i2v = upload(pick(R.c[1])); P.vvd = upload(pick(R.c[2]));
		P.vv2 = upload(pick(R.c[3])); P.vj1 = upload(pick(R.c[4])); P.vj2 = upload(pick(R.c[5]));
		P.ib0 = upload(pick(R.c[6])); P.b1 = upload(pick(R.c[7])); P.b2 = upload(pick(R.c[8]));
		P.Vd = dalloc<float>((size_t)3 * cnt); P.J = dalloc<float>((size_t)3 * cnt); P.Il = dalloc<float>(cnt);
		P.numTS = d_numTS;
		P.count = cnt;
		rlc_dev.push_back(P);
	}
	return 0;
}

// ------------------------------------------------------------------------------ finalize
int Engine::finalize()
{
	if (finalized) return fail("engine already finalized");
	if (!have_dense && !have_compressed) return fail("finalize: no operator uploaded");
	CK(cudaSetDevice(device));
	if (!stream) CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

	StageTimer tm("finalize", stream);
	std::vector<uint32_t> index32;
	if (have_dense) {
		if (compress_dense(index32)) return 1;
	}
	if (build_tables_and_index(index32)) return 1;
	std::vector<uint32_t>().swap(index32);
	tm.lap("tables");

	const size_t nfield = (size_t)3 * comp;
	d_V = dalloc<float>(nfield);
	d_I = dalloc<float>(nfield);
	if (!d_V || !d_I) return fail("out of device memory (fields)");
	d_numTS = dalloc<unsigned>(1);
	d_energy = dalloc<double>(2);
	d_flagE = dalloc<unsigned>(FLAG_WORDS);
	d_flagH = d_flagE + 1;
	d_halo_cnt = d_flagE + 2;
	d_halo_err = d_flagE + 4;
	if (sig_len) {
		d_sig[0] = upload(h_sig[0]);
		d_sig[1] = upload(h_sig[1]);
	}

	// stencil parameter blocks
	StencilParams base;
	memset(&base, 0, sizeof(base));
	base.V = d_V; base.I = d_I; base.idx = d_idx;
	base.nx = (int)gn[0]; base.ny = (int)gn[1]; base.nz = nzl;
	base.pitch = pitch; base.plane = plane; base.comp = comp;
	base.zchunk = tune_zchunk > 0 ? tune_zchunk : auto_zchunk();
	pE = base; pH = base;
	pE.tA = d_tab[0]; pE.tB = d_tab[1]; pE.tP0 = d_tab[2]; pE.tP1 = d_tab[3]; pE.tP2 = d_tab[4];
	pH.tA = d_tab[5]; pH.tB = d_tab[6]; pH.tP0 = d_tab[7]; pH.tP1 = d_tab[8]; pH.tP2 = d_tab[9];
	// update ranges in local planes: E on all owned planes, H on owned planes below the global top
	pE.k0 = (int)zb - z0; pE.k1 = (int)ze - z0;
	pH.k0 = (int)zb - z0; pH.k1 = (int)std::min(ze, gn[2] - 1) - z0;

	tm.lap("field set 0");
	if (build_pml()) return 1;
	tm.lap("upml boxes");
	pE.flux = d_flux_v; pH.flux = d_flux_i;
	sV[0] = d_V; sI[0] = d_I;
	// the one-pass schedule needs a second field set, no volume hooks between the half-steps and
	// disjoint UPML boxes (each is updated by its own shell launch)
	lor_fused = !h_lor.empty() && lorentz_fusable();
	fused_possible = fused_req != 0 && (h_lor.empty() || lor_fused) && h_rlc.empty() && pml_disjoint;
	if (!fused_possible) lor_fused = false;
	if (fused_possible) {
		sV[1] = dalloc<float>(nfield);
		sI[1] = dalloc<float>(nfield);
		if (!sV[1] || !sI[1]) {
			cudaGetLastError();
			fused_possible = false; // not enough memory for the ping-pong set: stay with two passes
		}
	}
	tm.lap("field set 1");
	if (build_mur()) return 1;
	if (build_exc()) return 1;
	if (build_lorentz()) return 1;
	if (build_rlc()) return 1;
	if (build_sheets()) return 1;
	if (build_tfsf()) return 1;
	ss_on = false;
	if (ss_period) {
		const unsigned cnt = (unsigned)ss_dir.size(); // may be 0 on a slab that owns none of the probes: energy only
		std::vector<long long> off(cnt);
		for (unsigned n = 0; n < cnt; ++n)
			off[n] = (long long)ss_dir[n] * comp + cell_off(ss_pos[n], ss_pos[(size_t)cnt + n], ss_pos[(size_t)2 * cnt + n]);
		memset(&pSs, 0, sizeof(pSs));
		pSs.V = d_V; pSs.I = d_I;
		if (off.empty()) off.push_back(0);
		pSs.off = upload(off);
		pSs.rec = dalloc<double>((size_t)2 * ss_period * std::max(1u, cnt));
		pSs.snap = dalloc<double>((size_t)2 * ss_period * std::max(1u, cnt));
		pSs.energy = dalloc<double>(4);
		pSs.info = dalloc<unsigned>(2);
		pSs.numTS = d_numTS;
		pSs.period = ss_period; pSs.count = cnt;
		pSs.nx = (int)gn[0]; pSs.ny = (int)gn[1];
		pSs.k0 = (int)zb - z0; pSs.k1 = (int)std::min(ze, gn[2] - 1) - z0;
		pSs.pitch = pitch; pSs.plane = plane; pSs.comp = comp;
		if (!pSs.off || !pSs.rec || !pSs.snap || !pSs.energy || !pSs.info) return fail("out of device memory (steady state)");
		ss_on = true;
	}
	if (fused_possible && build_fix_list()) return 1;
	if (fused_possible && lor_fused) {
		// k_fix_H recomputes plain H cells: a dispersive H cell next to a source / Mur plane keeps the two-pass schedule
		for (long long q = 0; q < fix_count && lor_fused; ++q) {
			const int* c = &h_fix_cells[3 * (size_t)q];
			for (const LorDev& D : lor_dev) {
				if (!D.i_on) continue;
				const int2 R = D.h_rows[(size_t)c[2] * gn[1] + c[1]];
				if ((unsigned)(c[0] - (R.x & 0xffff)) < (unsigned)(R.x >> 16)) lor_fused = false;
			}
		}
		if (!lor_fused) fused_possible = false;
	}
	CK(cudaStreamSynchronize(stream));
	CK(cudaGetLastError());
	tm.lap("hooks, fix list");

	build_schedule();
	tm.lap("schedule");
	finalized = true;
	// free host staging
	std::vector<oems_coeff_entry>().swap(h_table);
	for (auto& B : h_upml) for (int w = 0; w < 6; ++w) std::vector<float>().swap(B.c[w]);
	return 0;
}

// ------------------------------------------------------------------------------ schedule
template <typename K, typename P> static void launch1d(K kern, const P& p, long long n, cudaStream_t s)
{
	if (n <= 0) return;
	const int bs = 128;
	launch_k(kern, (unsigned)((n + bs - 1) / bs), bs, 0, s, p);
}

void Engine::build_schedule()
{
	step.clear();
	labels.clear();
	// one-pass schedule when it is possible (second field set allocated, no Lorentz/RLC hooks, UPML edge
	// path not in use) and either requested or -- automatic choice -- the mesh is big enough to fill the
	// GPU with z-marching blocks: below ~160^3 cells the two-pass kernels, which have a thread per cell
	// column and z chunk, are faster (tools/size_sweep.py, profiles/experiments_r01.md #13)
	fused_active = fused_possible && !edge_dirty && (fused_req == 1 || (fused_req < 0 && fused_auto_choice())) && !(lor_fused && !tma_req);
	// two-pass kernels with one cell per thread below 300 M cells per GPU (faster than the float4 z-march kernels up
	// to 640^3, equal at 1024 x 1024 x 512; option "small": 1 / 0 force, -1 automatic)
	small_active = small_req > 0 || (small_req < 0 && (long long)gn[0] * gn[1] * (ze - zb) < small_max_cells);
	const bool i16 = index_bytes == 2;
	const dim3 block(32, tune_rows);
	auto stencil_grid = [&](const StencilParams& p, int rows_total) {
		return dim3((unsigned)((pitch / 4 + 31) / 32), (unsigned)((rows_total + tune_rows - 1) / tune_rows),
		            (unsigned)std::max(1, (p.k1 - p.k0 + p.zchunk - 1) / p.zchunk));
	};
	const bool multi = peers_linked;

	// ---- pre-voltage hooks, reverse priority order (engine.cpp:224-230):
	//      Mur, Lorentz, RLC  (UPML pre is fused into the E kernel; Excitation has no pre hook)
	if (pMur.nplanes) (labels.push_back("mur_pre"), step.push_back([this](cudaStream_t s) { launch1d(k_mur_pre, pMur, pMur.total, s); }));
	// list order: [.., RLC, CondSheet, Lorentz, Mur, Excitation]; pre-hooks walk it
	// back to front: Mur, Lorentz, RLC
	for (size_t o = 0; o < lor_dev.size(); ++o)
		if (lor_dev[o].v_on) (labels.push_back("lorentz_pre_V"), step.push_back([this, o](cudaStream_t s) { launch1d(k_lorentz_pre, lor_dev[o].v, lor_dev[o].v.count, s); }));
	for (size_t r = 0; r < rlc_dev.size(); ++r)
		(labels.push_back("rlc_pre"), step.push_back([this, r](cudaStream_t s) { launch1d(k_rlc_pre, rlc_dev[r], rlc_dev[r].count, s); }));
	// absorbing sheets were inserted last (openems.cpp:1242-1243): first in the apply list, last here
	for (size_t a = 0; a < sheet_dev.size(); ++a)
		(labels.push_back("sheet_pre_V"), step.push_back([this, a](cudaStream_t s) { launch1d(k_sheet_pre, sheet_dev[a].v, sheet_dev[a].v.count, s); }));
	// ---- multi-GPU: the ghost H plane of this step must have arrived
	if (multi && peer_lo)
		(labels.push_back("halo_wait_H"), step.push_back([this](cudaStream_t s) {
			WaitParams w{d_flagH, d_numTS, 0u, d_halo_err, halo_timeout_cycles()};
			launch_k(k_halo_wait, 1, 1, 0, s, w);
		}));
	// ---- E half-step with fused UPML
	if (pE.k1 > pE.k0)
		(labels.push_back("update_E"), step.push_back([this, i16, block, stencil_grid](cudaStream_t s) {
			if (small_active) { // one cell per thread (small meshes)
				const dim3 g((unsigned)((gn[0] + 31) / 32), (unsigned)((pE.ny + tune_rows - 1) / tune_rows), (unsigned)(pE.k1 - pE.k0));
				if (i16) { if (has_pml) launch_k(k_small_E<uint16_t, true>, g, block, 0, s, pE); else launch_k(k_small_E<uint16_t, false>, g, block, 0, s, pE); }
				else { if (has_pml) launch_k(k_small_E<uint32_t, true>, g, block, 0, s, pE); else launch_k(k_small_E<uint32_t, false>, g, block, 0, s, pE); }
				return;
			}
			const dim3 g = stencil_grid(pE, pE.ny);
			if (i16) { if (has_pml) launch_k(k_update_E<uint16_t, true>, g, block, 0, s, pE); else launch_k(k_update_E<uint16_t, false>, g, block, 0, s, pE); }
			else { if (has_pml) launch_k(k_update_E<uint32_t, true>, g, block, 0, s, pE); else launch_k(k_update_E<uint32_t, false>, g, block, 0, s, pE); }
		}));
	// ---- post-voltage hooks: UPML (fused), TFSF, absorbing sheets, Mur
	if (pTfsf[0].groups) (labels.push_back("tfsf_V"), step.push_back([this](cudaStream_t s) { launch1d(k_tfsf, pTfsf[0], pTfsf[0].groups, s); }));
	for (size_t a = sheet_dev.size(); a-- > 0;)
		(labels.push_back("sheet_post_V"), step.push_back([this, a](cudaStream_t s) { launch1d(k_sheet_post, sheet_dev[a].v, sheet_dev[a].v.count, s); }));
	if (pMur.nplanes) (labels.push_back("mur_post"), step.push_back([this](cudaStream_t s) { launch1d(k_mur_post, pMur, pMur.total, s); }));
	// ---- apply-voltage hooks in list order: SteadyState, RLC, Lorentz, Mur, Excitation
	auto ss_launch = [this](const SsParams& q, cudaStream_t s) {
		launch_k(k_ss_record, std::max(1u, (q.count + 63) / 64), 64, 0, s, q);
		launch_k(k_ss_energy, 148 * 2, dim3(32, 8), 0, s, q);
		launch_k(k_ss_snapshot, 8, 256, 0, s, q);
	};
	if (ss_on) (labels.push_back("steadystate"), step.push_back([this, ss_launch](cudaStream_t s) { ss_launch(pSs, s); }));
	for (size_t a = sheet_dev.size(); a-- > 0;)
		(labels.push_back("sheet_apply_V"), step.push_back([this, a](cudaStream_t s) { launch1d(k_sheet_apply_V, sheet_dev[a].v, sheet_dev[a].v.count, s); }));
	for (size_t r = rlc_dev.size(); r-- > 0;) // same priority: reversed insertion order
		(labels.push_back("rlc_apply"), step.push_back([this, r](cudaStream_t s) { launch1d(k_rlc_apply, rlc_dev[r], rlc_dev[r].count, s); }));
	for (size_t o = 0; o < lor_dev.size(); ++o)
		if (lor_dev[o].v_on) (labels.push_back("lorentz_apply_V"), step.push_back([this, o](cudaStream_t s) { launch1d(k_lorentz_apply, lor_dev[o].v, lor_dev[o].v.count, s); }));
	const bool mur_exc = pMur.nplanes && pMur.total > 0 && pExc[0].groups && mur_exc_disjoint();
	if (mur_exc)
		(labels.push_back("mur_apply+excite_V"), step.push_back([this](cudaStream_t s) {
			const unsigned mb = (unsigned)((pMur.total + 127) / 128), eb = (pExc[0].groups + 127) / 128;
			launch_k(k_mur_apply_excite, mb + eb, 128, 0, s, pMur, pExc[0], mb);
		}));
	else if (pMur.nplanes) (labels.push_back("mur_apply"), step.push_back([this](cudaStream_t s) { launch1d(k_mur_apply, pMur, pMur.total, s); }));
	if (pExc[0].groups && !mur_exc) (labels.push_back("excite_V"), step.push_back([this](cudaStream_t s) { launch1d(k_excite, pExc[0], pExc[0].groups, s); }));
	// ---- multi-GPU: tangential E of my lowest owned plane -> lower neighbour's ghost plane
	if (multi && peer_lo)
		(labels.push_back("halo_push_E"), step.push_back([this](cudaStream_t s) {
			HaloParams h{d_V, peer_lo_V, (long long)((int)zb - z0) * plane, peer_lo_ghostE_off, comp, peer_lo_comp, plane,
			             d_halo_cnt, peer_lo_flagE, d_numTS, 1u};
			launch_k(k_halo_push, 296, 256, 0, s, h);
		}));
	// ---- pre-current hooks: Lorentz (UPML fused)
	for (size_t o = 0; o < lor_dev.size(); ++o)
		if (lor_dev[o].i_on) (labels.push_back("lorentz_pre_I"), step.push_back([this, o](cudaStream_t s) { launch1d(k_lorentz_pre, lor_dev[o].i, lor_dev[o].i.count, s); }));
	for (size_t a = 0; a < sheet_dev.size(); ++a)
		if (sheet_dev[a].i.count) (labels.push_back("sheet_pre_I"), step.push_back([this, a](cudaStream_t s) { launch1d(k_sheet_pre, sheet_dev[a].i, sheet_dev[a].i.count, s); }));
	if (multi && peer_hi)
		(labels.push_back("halo_wait_E"), step.push_back([this](cudaStream_t s) {
			WaitParams w{d_flagE, d_numTS, 1u, d_halo_err, halo_timeout_cycles()};
			launch_k(k_halo_wait, 1, 1, 0, s, w);
		}));
	// ---- H half-step with fused UPML, then the UPML cells the stencil never visits
	if (pH.k1 > pH.k0)
		(labels.push_back("update_H"), step.push_back([this, i16, block, stencil_grid](cudaStream_t s) {
			if (small_active) {
				const dim3 g((unsigned)((gn[0] - 1 + 31) / 32), (unsigned)((pH.ny - 1 + tune_rows - 1) / tune_rows), (unsigned)(pH.k1 - pH.k0));
				if (i16) { if (has_pml) launch_k(k_small_H<uint16_t, true>, g, block, 0, s, pH); else launch_k(k_small_H<uint16_t, false>, g, block, 0, s, pH); }
				else { if (has_pml) launch_k(k_small_H<uint32_t, true>, g, block, 0, s, pH); else launch_k(k_small_H<uint32_t, false>, g, block, 0, s, pH); }
				return;
			}
			const dim3 g = stencil_grid(pH, pH.ny - 1);
			if (i16) { if (has_pml) launch_k(k_update_H<uint16_t, true>, g, block, 0, s, pH); else launch_k(k_update_H<uint16_t, false>, g, block, 0, s, pH); }
			else { if (has_pml) launch_k(k_update_H<uint32_t, true>, g, block, 0, s, pH); else launch_k(k_update_H<uint32_t, false>, g, block, 0, s, pH); }
		}));
	if (pEdge.count && edge_dirty)
		(labels.push_back("upml_untouched_H"), step.push_back([this, i16](cudaStream_t s) {
			if (i16) launch1d(k_upml_untouched_H<uint16_t>, pEdge, pEdge.count, s);
			else launch1d(k_upml_untouched_H<uint32_t>, pEdge, pEdge.count, s);
		}));
	// ---- post-current hooks: UPML (fused), TFSF; then absorbing sheets (super-absorption), Lorentz, Excitation
	if (pTfsf[1].groups) (labels.push_back("tfsf_I"), step.push_back([this](cudaStream_t s) { launch1d(k_tfsf, pTfsf[1], pTfsf[1].groups, s); }));
	for (size_t a = sheet_dev.size(); a-- > 0;)
		if (sheet_dev[a].i.count) (labels.push_back("sheet_post_I"), step.push_back([this, a](cudaStream_t s) { launch1d(k_sheet_post, sheet_dev[a].i, sheet_dev[a].i.count, s); }));
	for (size_t a = sheet_dev.size(); a-- > 0;)
		if (sheet_dev[a].i.count) (labels.push_back("sheet_apply_I"), step.push_back([this, a](cudaStream_t s) { launch1d(k_sheet_apply_I, sheet_dev[a].i, sheet_dev[a].i.count, s); }));
	for (size_t o = 0; o < lor_dev.size(); ++o)
		if (lor_dev[o].i_on) (labels.push_back("lorentz_apply_I"), step.push_back([this, o](cudaStream_t s) { launch1d(k_lorentz_apply, lor_dev[o].i, lor_dev[o].i.count, s); }));
	if (pExc[1].groups) (labels.push_back("excite_I"), step.push_back([this](cudaStream_t s) { launch1d(k_excite, pExc[1], pExc[1].groups, s); }));
	if (multi && peer_hi)
		(labels.push_back("halo_push_H"), step.push_back([this](cudaStream_t s) {
			HaloParams h{d_I, peer_hi_I, (long long)((int)ze - 1 - z0) * plane, peer_hi_ghostH_off, comp, peer_hi_comp, plane,
			             d_halo_cnt + 1, peer_hi_flagH, d_numTS, 1u};
			launch_k(k_halo_push, 296, 256, 0, s, h);
		}));
	// the timestep counter: advanced by k_small_H itself when that is the last kernel of the timestep
	pH.tick = nullptr;
	if (small_active && !labels.empty() && labels.back() == "update_H" && pH.k1 > pH.k0) pH.tick = d_numTS;
	else (labels.push_back("tick"), step.push_back([this](cudaStream_t s) { launch_k(k_tick, 1, 1, 0, s, d_numTS); }));
	kernels_per_step = (unsigned)step.size();
	if (fused_active) build_schedule_fused();

	// ---- capture one timestep into a CUDA graph (launch-bound small meshes)
	if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
	if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
	for (int q = 0; q < 2; ++q) {
		if (graphf_exec[q]) { cudaGraphExecDestroy(graphf_exec[q]); graphf_exec[q] = nullptr; }
		if (graphf[q]) { cudaGraphDestroy(graphf[q]); graphf[q] = nullptr; }
	}
	use_graph = tune_graph != 0;
	auto capture = [&](std::vector<std::function<void(cudaStream_t)>>& list, cudaGraph_t& g, cudaGraphExec_t& ge) {
		if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
		for (auto& f : list) f(stream);
		if (cudaStreamEndCapture(stream, &g) != cudaSuccess || !g || cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
			cudaGetLastError();
			return false;
		}
		return true;
	};
	if (use_graph) {
		if (fused_active) use_graph = capture(stepf[0], graphf[0], graphf_exec[0]) && capture(stepf[1], graphf[1], graphf_exec[1]);
		else use_graph = capture(step, graph, graph_exec);
	}
}

// ---------------------------------------------------------------------------- fused schedule
// H cells whose curl reads an E value that a hook changes after the fused kernel ran:
// Mur planes (Apply2Voltages overwrites the boundary cells) and soft/hard E sources.
int Engine::build_fix_list()
{
	std::vector<long long> keys; // ((z*ny)+y)*nx+x, global
	const long long nx = gn[0], ny = gn[1];
	auto add = [&](int n, long long x, long long y, long long z) {
		// H cells that read V_n(x,y,z): engine.cpp:179-221
		const int d1[3][3] = {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}}; // V0: p-z, V1: p-x, V2: p-y  (first partner)
		const int d2[3][3] = {{0, 1, 0}, {0, 0, 1}, {1, 0, 0}}; // V0: p-y, V1: p-z, V2: p-x  (second partner)
		const long long c[3][3] = {{x, y, z}, {x - d1[n][0], y - d1[n][1], z - d1[n][2]}, {x - d2[n][0], y - d2[n][1], z - d2[n][2]}};
		for (int q = 0; q < 3; ++q) {
			const long long cx = c[q][0], cy = c[q][1], cz = c[q][2];
			if (cx < 0 || cy < 0 || cz < 0 || cx >= nx - 1 || cy >= ny - 1 || cz >= (long long)gn[2] - 1) continue;
			if (!owned((unsigned)cz)) continue;
			keys.push_back((cz * ny + cy) * nx + cx);
		}
	};
	for (const MurHost& M : h_mur) {
		const int nyP = (M.ny + 1) % 3, nyPP = (M.ny + 2) % 3;
		for (unsigned a = 0; a < M.n[0]; ++a)
			for (unsigned b = 0; b < M.n[1]; ++b) {
				long long pos[3];
				pos[M.ny] = M.line; pos[nyP] = a; pos[nyPP] = b;
				add(nyP, pos[0], pos[1], pos[2]);
				add(nyPP, pos[0], pos[1], pos[2]);
			}
	}
	if (h_tfsf.on) // DoPostVoltageUpdates adds to the tangential E on the six faces
		for (int n = 0; n < 3; ++n) {
			const int nP = (n + 1) % 3, nPP = (n + 2) % 3;
			for (int l = 0; l < 2; ++l) {
				if (!h_tfsf.active[n][l]) continue;
				long long pos[3];
				pos[n] = l ? h_tfsf.stop[n] : h_tfsf.start[n];
				for (unsigned a = h_tfsf.start[nP]; a <= h_tfsf.stop[nP]; ++a)
					for (unsigned b = h_tfsf.start[nPP]; b <= h_tfsf.stop[nPP]; ++b) {
						pos[nP] = a; pos[nPP] = b;
						add(nP, pos[0], pos[1], pos[2]);
						add(nPP, pos[0], pos[1], pos[2]);
					}
			}
		}
	for (const SheetHost& S : h_sheet) { // Apply2Voltages overwrites the two tangential components on the sheet
		const int nP = (S.ny + 1) % 3, nPP = (S.ny + 2) % 3;
		long long pos[3];
		pos[S.ny] = S.x0[S.ny];
		for (unsigned a = S.x0[nP]; a <= S.x1[nP]; ++a)
			for (unsigned b = S.x0[nPP]; b <= S.x1[nPP]; ++b) {
				pos[nP] = a; pos[nPP] = b;
				add(nP, pos[0], pos[1], pos[2]);
				add(nPP, pos[0], pos[1], pos[2]);
			}
	}
	const ExcHost& E = h_exc[0];
	for (size_t n = 0; n < E.dir.size(); ++n) add((int)E.dir[n], E.idx[0][n], E.idx[1][n], E.idx[2][n]);
	std::sort(keys.begin(), keys.end());
	keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
	fix_count = (long long)keys.size();
	std::vector<int> cells((size_t)3 * keys.size());
	for (size_t q = 0; q < keys.size(); ++q) {
		const long long k = keys[q];
		cells[3 * q] = (int)(k % nx);
		cells[3 * q + 1] = (int)((k / nx) % ny);
		cells[3 * q + 2] = (int)(k / (nx * ny)) - z0;
	}
	d_fix_cells = fix_count ? upload(cells) : nullptr;
	h_fix_cells = cells;
	if (fix_count && !d_fix_cells) return fail("out of device memory (fix-up list)");
	return 0;
}

// TMA descriptors of the source set of one parity: H (box 136 x 9 x 1 x 3), E (136 x 8 x 1 x 3) and
// the operator index (136 x 8 x 1); out-of-range elements are zero-filled (kernels_fused_tma.cuh)
int Engine::make_tma_maps(int par)
{
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
	                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static EncodeFn encode = nullptr;
	if (!encode) {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
			cudaGetLastError();
			return 1;
		}
		encode = (EncodeFn)fn;
	}
	{
		// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: set it on every
		// device an engine of this process lives on (two engines on two GPUs: link_engines_in_process)
		static std::vector<int> attr_devices;
		if (std::find(attr_devices.begin(), attr_devices.end(), device) == attr_devices.end()) {
			CK(cudaSetDevice(device));
			const int s16 = ft_smem_bytes<uint16_t, FT_STAGES>(), s32 = ft_smem_bytes<uint32_t, FT_STAGES>();
			cudaFuncSetAttribute(k_fused_tma<uint16_t, true, FT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16);
			cudaFuncSetAttribute(k_fused_tma<uint16_t, false, FT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16);
			cudaFuncSetAttribute(k_fused_tma<uint32_t, true, FT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, s32);
			cudaFuncSetAttribute(k_fused_tma<uint32_t, false, FT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, s32);
			cudaFuncSetAttribute(k_fused_tma<uint16_t, true, FT_STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16);
			cudaFuncSetAttribute(k_fused_tma<uint16_t, false, FT_STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16);
			cudaFuncSetAttribute(k_fused_tma<uint32_t, true, FT_STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s32);
			cudaFuncSetAttribute(k_fused_tma<uint32_t, false, FT_STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s32);
			if (cudaGetLastError() != cudaSuccess) return 1;
			attr_devices.push_back(device);
		}
	}
	const int S = par;
	const cuuint64_t dim4[4] = {(cuuint64_t)pitch, (cuuint64_t)gn[1], (cuuint64_t)nzl, 3};
	const cuuint64_t str4[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)plane * 4, (cuuint64_t)comp * 4};
	const cuuint32_t ones[4] = {1, 1, 1, 1};
	const cuuint32_t boxI[4] = {FT_W, FT_ROWS_I, 1, 3}, boxV[4] = {FT_W, FT_ROWS_V, 1, 3}, boxX[3] = {FT_W, FT_ROWS_V, 1};
	const cuuint64_t dim3[3] = {(cuuint64_t)pitch, (cuuint64_t)gn[1], (cuuint64_t)nzl};
	const cuuint64_t str3[2] = {(cuuint64_t)pitch * index_bytes, (cuuint64_t)plane * index_bytes};
	FusedTmaParams& T = pFT[par];
	CUresult r = encode(&T.mI, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, sI[S], dim4, str4, boxI, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r == CUDA_SUCCESS)
		r = encode(&T.mV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, sV[S], dim4, str4, boxV, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
		           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r == CUDA_SUCCESS)
		r = encode(&T.mX, index_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d_idx, dim3, str3, boxX,
		           ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS ? 0 : 1;
}

// TMA descriptors of k_xslab_tma (kernels_xslab_tma.cuh): same tensors as make_tma_maps, boxes of 24 columns
int Engine::make_xslab_maps(int par)
{
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
	                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
		cudaGetLastError();
		return 1;
	}
	EncodeFn encode = (EncodeFn)fn;
	{
		static std::vector<int> attr_devices;
		if (std::find(attr_devices.begin(), attr_devices.end(), device) == attr_devices.end()) {
			cudaSetDevice(device);
			cudaFuncSetAttribute(k_xslab_tma<uint16_t, XT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, xt_smem_bytes<uint16_t, XT_STAGES>());
			cudaFuncSetAttribute(k_xslab_tma<uint32_t, XT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, xt_smem_bytes<uint32_t, XT_STAGES>());
			if (cudaGetLastError() != cudaSuccess) return 1;
			attr_devices.push_back(device);
		}
	}
	const int S = par;
	XTmaParams& T = pXt[par];
	const cuuint64_t dim4[4] = {(cuuint64_t)pitch, (cuuint64_t)gn[1], (cuuint64_t)nzl, 3};
	const cuuint64_t str4[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)plane * 4, (cuuint64_t)comp * 4};
	const cuuint32_t ones[4] = {1, 1, 1, 1};
	const cuuint32_t boxI[4] = {XT_W, XT_ROWS_I, 1, 3}, boxV[4] = {XT_W, XT_ROWS_V, 1, 3}, boxX[3] = {XT_W, XT_ROWS_V, 1};
	const cuuint64_t dim3[3] = {(cuuint64_t)pitch, (cuuint64_t)gn[1], (cuuint64_t)nzl};
	const cuuint64_t str3[2] = {(cuuint64_t)pitch * index_bytes, (cuuint64_t)plane * index_bytes};
	CUresult r = encode(&T.mI, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, sI[S], dim4, str4, boxI, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r == CUDA_SUCCESS)
		r = encode(&T.mV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, sV[S], dim4, str4, boxV, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
		           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r == CUDA_SUCCESS)
		r = encode(&T.mX, index_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d_idx, dim3, str3, boxX,
		           ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
		           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS ? 0 : 1;
}

void Engine::build_schedule_fused()
{
	const bool i16 = index_bytes == 2;
	const bool multi = peers_linked;
	labelsf.clear();
	sched_error.clear();
	if (multi && !side_stream) {
		if (cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); side_stream = nullptr; }
		for (int q = 0; q < 2 && side_stream; ++q) {
			cudaEventCreateWithFlags(&ev_fork[q][0], cudaEventDisableTiming);
			cudaEventCreateWithFlags(&ev_fork[q][1], cudaEventDisableTiming);
			cudaEventCreateWithFlags(&ev_join[q], cudaEventDisableTiming);
		}
	}
	// TMA descriptors of both source sets; without them the register-staged kernel is used
	tma_active = tma_req != 0 && make_tma_maps(0) == 0 && make_tma_maps(1) == 0;
	// UPML boxes updated by their own one-pass kernel k_xslab_EH ("x slabs", kernels_xslab.cuh): thin in x, at
	// the low or the high end of the mesh, the float4 chunks they touch (+ the halo column on the low side)
	// fit a 16-line window, at most one per end, no H cell of the fix-up list inside (those are recomputed
	// from the hooks' final E, which an in-place flux cannot redo), and -- high end -- the box must not start on a
	// chunk boundary: the big kernel computes the first line of that chunk with the plain formula for its own
	// last H line
	xs_box[0] = xs_box[1] = -1;
	xs_win[0] = xs_win[1] = 0;
	xslab_tma = tma_active && xslab_req == 2;
	if (tma_active && xslab_req) {
		for (int b = 0; b < pE.nboxes; ++b) {
			const PmlBox& B = pE.box[b];
			int g = -1;
			if (xslab_tma) {
				// k_xslab_tma: a window of 16 lines starting on a 64-byte boundary must hold the box (+ the first line
				// right of it on the low side) and end at / beyond the last line on the high side
				// (whole 32-byte sectors: the window starts on a multiple of 8 lines)
				const int ws = B.s[0] / 8 * 8;
				if (B.s[0] == 0 && B.n[0] + 1 <= 16 && (int)gn[0] >= 48) { g = 0; xs_win[0] = 0; }
				else if (B.s[0] + B.n[0] == (int)gn[0] && ws >= 32 && ws + 16 >= (int)gn[0] && ws + 16 <= pitch) { g = 1; xs_win[1] = ws; }
				// plain cells of a window are updated by k_xslab_tma, which has no ADE branch
				if (g >= 0 && lor_fused && lor_xmax >= xs_win[g] - 1 && lor_xmin <= xs_win[g] + 16) g = -1;
			} else {
				if (B.s[0] == 0 && (B.n[0] + 3) / 4 * 4 + 1 <= 16 && (B.n[0] + 3) / 4 * 4 < (int)gn[0]) g = 0;
				else if (B.s[0] + B.n[0] == (int)gn[0] && B.s[0] % 4 != 0 && (int)gn[0] - B.s[0] / 4 * 4 <= 16 && B.s[0] / 4 * 4 > 0) g = 1;
			}
			if (g < 0 || xs_box[g] >= 0) continue;
			bool hit = false;
			for (long long q = 0; q < fix_count && !hit; ++q) {
				const int* c = &h_fix_cells[3 * (size_t)q];
				const int lo = xslab_tma ? xs_win[g] : B.s[0], n = xslab_tma ? 16 : B.n[0];
				hit = (unsigned)(c[0] - lo) < (unsigned)n && (unsigned)(c[1] - B.s[1]) < (unsigned)B.n[1] && (unsigned)(c[2] - B.s[2]) < (unsigned)B.n[2];
			}
			if (!hit) xs_box[g] = b;
		}
		// both windows in one mesh that is narrower than the two of them: keep the low one
		if (xs_box[0] >= 0 && xs_box[1] >= 0 && (pE.box[xs_box[0]].n[0] + 3) / 4 * 4 + 1 > pE.box[xs_box[1]].s[0] / 4 * 4) xs_box[1] = -1;
		if ((xs_box[0] >= 0 || xs_box[1] >= 0) && !d_flux_v2) {
			d_flux_v2 = dalloc<float>((size_t)flux_floats);
			if (!d_flux_v2) { cudaGetLastError(); xs_box[0] = xs_box[1] = -1; }
			else cudaMemsetAsync(d_flux_v2, 0, (size_t)flux_floats * sizeof(float), stream);
		}
	}
	for (int par = 0; par < 2; ++par) {
		const int S = par, D = par ^ 1;
		auto& L = stepf[par];
		L.clear();
		auto lab = [&](const char* name) { if (par == 0) labelsf.push_back(name); };
		// ---- parameter blocks of this parity
		FusedParams& F = pF[par];
		memset(&F, 0, sizeof(F));
		F.Vs = sV[S]; F.Is = sI[S]; F.Vd = sV[D]; F.Id = sI[D];
		F.idx = d_idx;
		F.eA = d_tab[0]; F.eB = d_tab[1];
		F.hA = d_tab[5]; F.hB = d_tab[6];
		F.nx = (int)gn[0]; F.ny = (int)gn[1]; F.nz = nzl;
		F.pitch = pitch; F.plane = plane; F.comp = comp;
		F.kE0 = pE.k0; F.kE1 = pE.k1;
		F.kH0 = pH.k0;
		F.kH1 = (multi && peer_hi) ? pE.k1 - 1 : pH.k1;   // the slab's top plane waits for the ghost E plane
		F.kHc1 = (multi && peer_hi) ? F.kH1 : pE.k1;       // planes above kH1 are copied through (top of the domain)
		F.zchunk = has_pml ? std::min(pE.zchunk, 63) : pE.zchunk; // the kernel keeps one shell bit per plane of a chunk
		// TMA-staged kernel: short marches keep the concurrently running blocks on the same few planes
		// (measured optimum 16..24 planes at 1024^3, profiles/experiments_r01.md)
		if (tma_req && tune_zchunk <= 0) {
			const long long tiles = (long long)((pitch / 4 + 31) / 32) * ((gn[1] + FUSED_TY - 1) / FUSED_TY);
			const int planes = std::max(1, F.kE1 - F.kE0);
			F.zchunk = (tiles * ((planes + 15) / 16) >= 27 * 296) ? 16 : 8; // mid-size meshes: more, shorter marches
		}
		// UPML shell: all boxes of a half-step in one launch (kernels_fused.cuh)
		ShellParams& SE = pShE[par];
		ShellParams& SH = pShH[par];
		memset(&SE, 0, sizeof(SE));
		SE.idx = d_idx;
		SE.nx = (int)gn[0]; SE.ny = (int)gn[1]; SE.pitch = pitch; SE.plane = plane; SE.comp = comp;
		SH = SE;
		SE.Xs = sV[S]; SE.Xd = sV[S]; SE.Y = sI[S]; // in place: see kernels_fused.cuh
		SE.tA = d_tab[0]; SE.tB = d_tab[1]; SE.tP0 = d_tab[2]; SE.tP1 = d_tab[3]; SE.tP2 = d_tab[4];
		SH.Xs = sI[S]; SH.Xd = sI[D]; SH.Y = sV[D];
		SH.tA = d_tab[5]; SH.tB = d_tab[6]; SH.tP0 = d_tab[7]; SH.tP1 = d_tab[8]; SH.tP2 = d_tab[9];
		int ns = 0;
		float* const fluxV[2] = {d_flux_v, d_flux_v2};
		FusedTmaParams& FT = pFT[par];
		memset(FT.xs, 0, sizeof(FT.xs));
		XSlabParams& XP = pXs[par];
		memset(&XP, 0, sizeof(XP));
		XP.Vs = sV[S]; XP.Is = sI[S]; XP.Vd = sV[D]; XP.Id = sI[D];
		XP.idx = d_idx;
		XP.eA = d_tab[0]; XP.eB = d_tab[1]; XP.eP0 = d_tab[2]; XP.eP1 = d_tab[3]; XP.eP2 = d_tab[4];
		XP.hA = d_tab[5]; XP.hB = d_tab[6]; XP.hP0 = d_tab[7]; XP.hP1 = d_tab[8]; XP.hP2 = d_tab[9];
		XP.nx = (int)gn[0]; XP.ny = (int)gn[1]; XP.nz = nzl;
		XP.pitch = pitch; XP.plane = plane; XP.comp = comp;
		XP.kE0 = F.kE0; XP.kE1 = F.kE1; XP.kH1 = F.kH1; XP.kHc1 = F.kHc1;
		XP.zchunk = xslab_tma ? xt_zchunk : 16;
		nxs = 0;
		const int pH_k1 = F.kH1;
		// ---- whole planes / whole rows at the ends of the mesh that consist of UPML cells only (the z-low / z-high
		// boxes together with the parts of the x and y boxes next to them; the y boxes with the x boxes' ends): the
		// one-pass kernel (and the x-window kernel) skip them altogether instead of passing them through --
		// k_shell_E stores E_new of those cells in both field sets (the source set in place for the neighbours'
		// curls, the destination set as the result), k_shell_H stores their H_new.  Saves 50 - 12 bytes per cell.
		F.jb = 0; F.je = (int)gn[1];
		SE.sk0 = F.kE0; SE.sk1 = F.kE1; SE.sjb = 0; SE.sje = (int)gn[1]; SE.Xd2 = nullptr;
		if (par == 0) skip_active = 0;
		if (tma_active && skip_req && pE.nboxes && xslab_req != 1) {
			const int nx = (int)gn[0], ny = (int)gn[1];
			// cells of all UPML boxes inside a region (the boxes are disjoint)
			auto covered = [&](int j0, int j1, int k0, int k1) {
				long long c = 0;
				for (int q = 0; q < pE.nboxes; ++q) {
					const PmlBox& B = pE.box[q];
					const long long dj = std::min(j1, B.s[1] + B.n[1]) - std::max(j0, B.s[1]);
					const long long dk = std::min(k1, B.s[2] + B.n[2]) - std::max(k0, B.s[2]);
					if (dj > 0 && dk > 0) c += dj * dk * B.n[0];
				}
				return c == (long long)(j1 - j0) * (k1 - k0) * nx;
			};
			int k0 = F.kE0, k1 = F.kE1, jb = 0, je = ny;
			for (int q = 0; q < pE.nboxes; ++q) { // thinnest box that starts at the bottom / ends at the top
				const PmlBox& B = pE.box[q];
				if (B.s[2] == F.kE0 && B.s[2] + B.n[2] < F.kE1 && covered(0, ny, F.kE0, B.s[2] + B.n[2])) k0 = std::max(k0, B.s[2] + B.n[2]);
				if (B.s[2] + B.n[2] == F.kE1 && B.s[2] > F.kE0 && covered(0, ny, B.s[2], F.kE1)) k1 = std::min(k1, B.s[2]);
			}
			if (k1 - k0 < 2) { k0 = F.kE0; k1 = F.kE1; }
			for (int q = 0; q < pE.nboxes; ++q) {
				const PmlBox& B = pE.box[q];
				if (B.s[1] == 0 && B.n[1] < ny && covered(0, B.n[1], k0, k1)) jb = std::max(jb, B.n[1]);
				if (B.s[1] + B.n[1] == ny && B.s[1] > 0 && covered(B.s[1], ny, k0, k1)) je = std::min(je, B.s[1]);
			}
			if (je - jb < 2) { jb = 0; je = ny; }
			if (k0 > F.kE0 || k1 < F.kE1 || jb > 0 || je < ny) {
				SE.sk0 = k0; SE.sk1 = k1; SE.sjb = jb; SE.sje = je; SE.Xd2 = sV[D];
				F.kE0 = k0; F.kE1 = k1; F.jb = jb; F.je = je;
				F.kH1 = std::min(F.kH1, k1); F.kHc1 = std::min(F.kHc1, k1);
				XP.kE0 = F.kE0; XP.kE1 = F.kE1; XP.kH1 = F.kH1; XP.kHc1 = F.kHc1;
				skip_active = (k0 > pE.k0) + (k1 < pE.k1) + (jb > 0) + (je < ny);
			}
		}
		// one entry of the two shell launches: rows [jr0, jr1) (box-local) and local planes [k0, k1) of box B
		int nent = 0;
		int xs_sh[2] = {-1, -1};
		auto add_shell = [&](const PmlBox& B, int jr0, int jr1, int k0, int k1, bool pingpong = false) {
			if (jr1 <= jr0 || k1 <= k0 || nent >= OEMS_MAX_SHELL_ENTRIES) return;
			ShellBoxParams q;
			memset(&q, 0, sizeof(q));
			q.cs = (long long)B.n[0] * B.n[1] * B.n[2];
			q.bs0 = B.s[0]; q.bs1 = B.s[1]; q.bs2 = B.s[2]; q.bn0 = B.n[0]; q.bn1 = B.n[1];
			q.jr0 = jr0; q.jr1 = jr1;
			q.c0 = B.s[0] / 4; q.nchunk = (B.s[0] + B.n[0] - 1) / 4 - q.c0 + 1;
			// lanes side by side in x: the smallest power of two that covers the box
			q.xl = q.nchunk <= 4 ? 4 : q.nchunk <= 8 ? 8 : q.nchunk <= 16 ? 16 : 32;
			const int rows = 8 * (32 / q.xl);
			q.gx = (q.nchunk + q.xl - 1) / q.xl;
			q.gy = (jr1 - jr0 + rows - 1) / rows;
			ShellBoxParams e = q, h = q;
			e.flux = (pingpong ? fluxV[S] : d_flux_v) + B.off;
			e.flux_out = (pingpong ? fluxV[D] : d_flux_v) + B.off;
			e.k0 = k0; e.k1 = k1;
			h.flux = h.flux_out = d_flux_i + B.off;
			h.k0 = std::max(k0, F.kH0); h.k1 = std::min(k1, pH_k1);
			// z chunk: long marches save the re-read of the carried plane, short ones give more blocks
			for (ShellBoxParams* w : {&e, &h}) {
				const int nk = std::max(0, w->k1 - w->k0);
				int zc = shell_zchunk;
				while (zc > 4 && (long long)w->gx * w->gy * ((nk + zc - 1) / zc) < 4 * 148) zc /= 2;
				w->zchunk = std::max(1, std::min(zc, nk));
			}
			SE.box[nent] = e;
			SH.box[nent] = h;
			++nent;
		};
		for (int b = 0; b < pE.nboxes; ++b) {
			const PmlBox& B = pE.box[b];
			if (b == xs_box[0] || b == xs_box[1]) {
				const int g = b == xs_box[0] ? 0 : 1;
				const int c0 = B.s[0] / 4, c1 = (B.s[0] + B.n[0] - 1) / 4; // chunks the box touches
				XSlabFoot& X = FT.xs[g];
				X.on = xslab_tma ? 0 : 1; X.c0 = c0; X.cn = c1 - c0 + 1; // k_xslab_tma overwrites its window after the big kernel: no store masks
				X.j0 = B.s[1]; X.jn = B.n[1]; X.k0 = B.s[2]; X.kn = B.n[2];
				X.x0 = B.s[0]; X.x1 = B.s[0] + B.n[0];
				XSlabBox& Q = XP.box[nxs++];
				Q.w0 = xslab_tma ? xs_win[g] : c0 * 4; Q.own0 = c0 * 4; Q.own1 = std::min((c1 + 1) * 4, (int)gn[0]);
				Q.bs0 = B.s[0]; Q.bn0 = B.n[0];
				Q.s1 = B.s[1]; Q.n1 = B.n[1]; Q.s2 = B.s[2]; Q.n2 = B.n[2];
				Q.cs = (long long)B.n[0] * B.n[1] * B.n[2];
				Q.fVs = fluxV[S] + B.off; Q.fVd = fluxV[D] + B.off; Q.fI = d_flux_i + B.off;
				// the window kernel owns the box cells on the rows / planes the one-pass kernels work on; the pieces in
				// skipped planes / rows go through the shell launches like the other boxes' cells there
				Q.oj0 = std::max(B.s[1], F.jb); Q.oj1 = std::min(B.s[1] + B.n[1], F.je);
				Q.ok0 = std::max(B.s[2], F.kE0); Q.ok1 = std::min(B.s[2] + B.n[2], F.kE1);
				if (xslab_tma && ns < OEMS_MAX_PML_BOXES) {
					// the big kernel need not compute the window cells (graded UPML coefficients: a table gather per
					// cell): it passes them through like shell cells and k_xslab_tma overwrites them.  Not the first
					// chunk of the high window: the plain cell left of the window takes its E_new from there.
					const int wc0 = Q.w0 / 4 + (g == 1 ? 1 : 0), wcn = 4 - (g == 1 ? 1 : 0);
					F.sh[ns].c0 = wc0; F.sh[ns].cn = wcn;
					F.sh[ns].j0 = Q.oj0; F.sh[ns].jn = Q.oj1 - Q.oj0;
					F.sh[ns].k0 = Q.ok0; F.sh[ns].kn = Q.ok1 - Q.ok0;
					xs_sh[g] = ns;
					++ns;
				}
				if (xslab_tma) {
					add_shell(B, 0, B.n[1], B.s[2], Q.ok0, true);
					add_shell(B, 0, B.n[1], Q.ok1, B.s[2] + B.n[2], true);
					add_shell(B, 0, Q.oj0 - B.s[1], Q.ok0, Q.ok1, true);
					add_shell(B, Q.oj1 - B.s[1], B.n[1], Q.ok0, Q.ok1, true);
				}
				continue;
			}
			F.sh[ns].c0 = B.s[0] / 4; F.sh[ns].cn = (B.s[0] + B.n[0] - 1) / 4 - F.sh[ns].c0 + 1;
			F.sh[ns].j0 = B.s[1]; F.sh[ns].jn = B.n[1];
			F.sh[ns].k0 = B.s[2]; F.sh[ns].kn = B.n[2];
			++ns;
			add_shell(B, 0, B.n[1], B.s[2], B.s[2] + B.n[2]);
		}
		F.nsh = ns;
		SE.nboxes = SH.nboxes = nent;
		XP.jb = F.jb; XP.je = F.je;
		if (tma_active && tune_zchunk <= 0) {
			// marches of equal length: 131 planes of a slab as 8 x 17 rather than 8 x 16 + 3 (every block pays the same
			// start-up: barrier set-up, first TMA round trip, the carried plane)
			auto even = [](int planes, int zc) {
				const int n = std::max(1, (planes + zc / 2) / zc);
				return std::min(63, (planes + n - 1) / n);
			};
			const int planes = std::max(1, F.kE1 - F.kE0);
			F.zchunk = even(planes, F.zchunk);
			if (xslab_tma) XP.zchunk = even(planes, XP.zchunk);
		}
		{   // the window kernel's list of foreign footprints: the boxes on the shell path, not its own windows
			int m = 0;
			for (int q = 0; q < ns; ++q) {
				if (q == xs_sh[0] || q == xs_sh[1]) continue;
				XP.sh[m].c0 = F.sh[q].c0; XP.sh[m].cn = F.sh[q].cn; XP.sh[m].j0 = F.sh[q].j0; XP.sh[m].jn = F.sh[q].jn; XP.sh[m].k0 = F.sh[q].k0; XP.sh[m].kn = F.sh[q].kn;
				++m;
			}
			XP.nsh = m;
		}
		for (ShellParams* w : {&SE, &SH}) {
			unsigned nb = 0;
			for (int b = 0; b < w->nboxes; ++b) {
				ShellBoxParams& q = w->box[b];
				const int nk = std::max(0, q.k1 - q.k0);
				q.blk0 = nb;
				nb += (unsigned)q.gx * q.gy * ((nk + q.zchunk - 1) / q.zchunk);
			}
			w->nblocks = nb;
		}
		FT.f = F;
		memset(&FT.lor, 0, sizeof(FT.lor));
		if (lor_fused) {
			if (!tma_active) sched_error = "the one-pass schedule with Lorentz/Drude material needs the TMA-staged kernel (no tensor map could be made)";
			FT.lor.nord = (int)lor_dev.size();
			for (size_t o = 0; o < lor_dev.size(); ++o) {
				const LorDev& Ld = lor_dev[o];
				FT.lor.o[o].rows = Ld.d_rows;
				FT.lor.o[o].ade_v = Ld.v_on ? Ld.v.ade : nullptr;
				FT.lor.o[o].ade_i = Ld.i_on ? Ld.i.ade : nullptr;
				FT.lor.o[o].count = Ld.v_on ? Ld.v.count : Ld.i.count;
			}
		}
		if (xslab_tma && nxs) {
			pXt[par].x = XP;
			if (make_xslab_maps(par)) { xslab_tma = false; }
		}
		FixParams& X = pFix[par];
		memset(&X, 0, sizeof(X));
		X.Is = sI[S]; X.Id = sI[D]; X.Vd = sV[D];
		X.idx = d_idx;
		X.hA = d_tab[5]; X.hB = d_tab[6];
		X.cell = d_fix_cells; X.count = fix_count;
		X.pitch = pitch; X.plane = plane; X.comp = comp;
		for (size_t a = 0; a < sheet_dev.size(); ++a) {
			// index 0: hooks that read the source set (pre), 1: hooks on the destination set (post, apply)
			pShV[par][a] = sheet_dev[a].v; pShI[par][a] = sheet_dev[a].i;
		}
		pMurS[par] = pMur; pMurS[par].V = sV[S];
		pMurD[par] = pMur; pMurD[par].V = sV[D];
		pExcD[par][0] = pExc[0]; pExcD[par][0].X = sV[D];
		pExcD[par][1] = pExc[1]; pExcD[par][1].X = sI[D];
		StencilParams& T = pHtop[par];
		T = pH;
		T.tick = nullptr;
		T.V = sV[D]; T.I = sI[S]; T.Iout = sI[D]; T.flux = d_flux_i; T.flux_out = nullptr; // flux in place
		T.k0 = pE.k1 - 1; T.k1 = pE.k1; T.zchunk = 1;

		// ---- pre-voltage hooks on the source set
		if (pMur.nplanes) { lab("mur_pre"); L.push_back([this, par](cudaStream_t s) { launch1d(k_mur_pre, pMurS[par], pMurS[par].total, s); }); }
		// Lorentz / Drude: both pre hooks (engine_ext_lorentzmaterial.cpp:79-127) read timestep-n values of the cell itself
		// -> list kernels on the source set; the one-pass kernel subtracts the advanced ADE values (LOR instance)
		if (lor_fused)
			for (size_t o = 0; o < lor_dev.size(); ++o) {
				if (lor_dev[o].v_on) {
					lab("lorentz_pre_V");
					L.push_back([this, par, o](cudaStream_t s) { LorParams q = lor_dev[o].v; q.X = sV[par]; launch1d(k_lorentz_pre, q, q.count, s); });
				}
				if (lor_dev[o].i_on) {
					lab("lorentz_pre_I");
					L.push_back([this, par, o](cudaStream_t s) { LorParams q = lor_dev[o].i; q.X = sI[par]; launch1d(k_lorentz_pre, q, q.count, s); });
				}
			}
		// absorbing sheets: both pre hooks read timestep-n values -> the source set, before k_shell_E touches it
		for (size_t a = 0; a < sheet_dev.size(); ++a) {
			lab("sheet_pre_V");
			L.push_back([this, par, a](cudaStream_t s) { SheetParams q = pShV[par][a]; q.X = sV[par]; launch1d(k_sheet_pre, q, q.count, s); });
			if (sheet_dev[a].i.count) {
				lab("sheet_pre_I");
				L.push_back([this, par, a](cudaStream_t s) { SheetParams q = pShI[par][a]; q.X = sI[par]; launch1d(k_sheet_pre, q, q.count, s); });
			}
		}
		if (multi && peer_lo) {
			lab("halo_wait_H");
			L.push_back([this](cudaStream_t s) {
				WaitParams w{d_flagH, d_numTS, 0u, d_halo_err, halo_timeout_cycles()};
				launch_k(k_halo_wait, 1, 1, 0, s, w);
			});
		}
		// ---- E of the UPML shell, then E and H of everything else in one pass
		if (has_pml) {
			lab("shell_E");
			L.push_back([this, par, i16](cudaStream_t s) {
				const ShellParams& q = pShE[par];
				if (!q.nblocks) return;
				if (i16) launch_k(k_shell_E<uint16_t>, q.nblocks, dim3(32, 8), 0, s, q); else launch_k(k_shell_E<uint32_t>, q.nblocks, dim3(32, 8), 0, s, q);
			});
		}
		if (nxs && !xslab_tma) {
			// x slabs: E and H in one pass, source -> destination set (after the shells of the other boxes: it
			// takes their E_new as neighbour values; before or after the big kernel makes no difference)
			lab("xslab_EH");
			L.push_back([this, par, i16](cudaStream_t s) {
				const XSlabParams& q = pXs[par];
				int rows = 0, planes = 0;
				for (int b = 0; b < nxs; ++b) { rows = std::max(rows, q.box[b].n1); planes = std::max(planes, q.box[b].n2); }
				const dim3 g((unsigned)((rows + XSLAB_ROWS - 1) / XSLAB_ROWS), (unsigned)((planes + q.zchunk - 1) / q.zchunk), (unsigned)nxs);
				if (i16) launch_k(k_xslab_EH<uint16_t>, g, dim3(16, 16), 0, s, q); else launch_k(k_xslab_EH<uint32_t>, g, dim3(16, 16), 0, s, q);
			});
		}
		lab("fused_EH");
		L.push_back([this, par, i16](cudaStream_t s) {
			const FusedParams& q = pF[par];
			const dim3 block(32, FUSED_TY + 1);
			const dim3 g((unsigned)((pitch / 4 + 31) / 32), (unsigned)((q.je - q.jb + FUSED_TY - 1) / FUSED_TY),
			             (unsigned)std::max(1, (q.kE1 - q.kE0 + q.zchunk - 1) / q.zchunk));
			if (tma_active) {
				const FusedTmaParams& t = pFT[par];
				const int sm = i16 ? ft_smem_bytes<uint16_t, FT_STAGES>() : ft_smem_bytes<uint32_t, FT_STAGES>();
				if (lor_fused) {
					if (i16) { if (has_pml) launch_k(k_fused_tma<uint16_t, true, FT_STAGES, true>, g, block, sm, s, t); else launch_k(k_fused_tma<uint16_t, false, FT_STAGES, true>, g, block, sm, s, t); }
					else { if (has_pml) launch_k(k_fused_tma<uint32_t, true, FT_STAGES, true>, g, block, sm, s, t); else launch_k(k_fused_tma<uint32_t, false, FT_STAGES, true>, g, block, sm, s, t); }
					return;
				}
				if (i16) { if (has_pml) launch_k(k_fused_tma<uint16_t, true, FT_STAGES>, g, block, sm, s, t); else launch_k(k_fused_tma<uint16_t, false, FT_STAGES>, g, block, sm, s, t); }
				else { if (has_pml) launch_k(k_fused_tma<uint32_t, true, FT_STAGES>, g, block, sm, s, t); else launch_k(k_fused_tma<uint32_t, false, FT_STAGES>, g, block, sm, s, t); }
				return;
			}
			if (i16) { if (has_pml) launch_k(k_fused_EH<uint16_t, true>, g, block, 0, s, q); else launch_k(k_fused_EH<uint16_t, false>, g, block, 0, s, q); }
			else { if (has_pml) launch_k(k_fused_EH<uint32_t, true>, g, block, 0, s, q); else launch_k(k_fused_EH<uint32_t, false>, g, block, 0, s, q); }
		});
		if (nxs && xslab_tma) {
			// x slabs, TMA-staged: overwrites its 16-line windows in the destination set AFTER the big kernel (which
			// treated them as plain cells) and before the hooks / k_shell_H, which read the final E of the window
			lab("xslab_EH");
			L.push_back([this, par, i16](cudaStream_t s) {
				const XTmaParams& q = pXt[par];
				const dim3 g((unsigned)((q.x.je - q.x.jb + XT_TY - 1) / XT_TY), (unsigned)std::max(1, (q.x.kE1 - q.x.kE0 + q.x.zchunk - 1) / q.x.zchunk), (unsigned)nxs);
				if (i16) launch_k(k_xslab_tma<uint16_t, XT_STAGES>, g, XT_THREADS, xt_smem_bytes<uint16_t, XT_STAGES>(), s, q);
				else launch_k(k_xslab_tma<uint32_t, XT_STAGES>, g, XT_THREADS, xt_smem_bytes<uint32_t, XT_STAGES>(), s, q);
			});
		}
		// ---- post / apply voltage hooks on the destination set
		pTfsfD[par][0] = pTfsf[0]; pTfsfD[par][0].X = sV[D];
		pTfsfD[par][1] = pTfsf[1]; pTfsfD[par][1].X = sI[D];
		if (pTfsf[0].groups) { lab("tfsf_V"); L.push_back([this, par](cudaStream_t s) { launch1d(k_tfsf, pTfsfD[par][0], pTfsfD[par][0].groups, s); }); }
		for (size_t a = sheet_dev.size(); a-- > 0;) {
			lab("sheet_post_V");
			L.push_back([this, par, a](cudaStream_t s) { SheetParams q = pShV[par][a]; q.X = sV[par ^ 1]; launch1d(k_sheet_post, q, q.count, s); });
		}
		if (pMur.nplanes) {
			lab("mur_post"); L.push_back([this, par](cudaStream_t s) { launch1d(k_mur_post, pMurD[par], pMurD[par].total, s); });
		}
		if (ss_on) {
			pSsF[par] = pSs; pSsF[par].V = sV[D]; pSsF[par].I = sI[S];
			lab("steadystate");
			L.push_back([this, par](cudaStream_t s) {
				const SsParams& q = pSsF[par];
				launch_k(k_ss_record, std::max(1u, (q.count + 63) / 64), 64, 0, s, q);
				launch_k(k_ss_energy, 148 * 2, dim3(32, 8), 0, s, q);
				launch_k(k_ss_snapshot, 8, 256, 0, s, q);
			});
		}
		for (size_t a = sheet_dev.size(); a-- > 0;) {
			lab("sheet_apply_V");
			L.push_back([this, par, a](cudaStream_t s) { SheetParams q = pShV[par][a]; q.X = sV[par ^ 1]; launch1d(k_sheet_apply_V, q, q.count, s); });
		}
		if (pMur.nplanes) {
			lab("mur_apply"); L.push_back([this, par](cudaStream_t s) { launch1d(k_mur_apply, pMurD[par], pMurD[par].total, s); });
		}
		if (pExc[0].groups) { lab("excite_V"); L.push_back([this, par](cudaStream_t s) { launch1d(k_excite, pExcD[par][0], pExcD[par][0].groups, s); }); }
		// Slabs: the E halo push, the wait for the upper neighbour's E plane and the H update of the slab's top plane
		// are small kernels with a long latency chain (NVLink, the neighbour's pace).  They run on a SIDE stream (a
		// parallel branch of the step graph) next to fix_H / k_shell_H and join before the current-side hooks: at 8
		// GPUs ~60 of the 1500 us of a timestep no longer sit at the end of the step with the GPU idle.
		const bool side = multi && overlap_halo && side_stream;
		if (multi && peer_lo) {
			lab("halo_push_E");
			L.push_back([this, D, par, side](cudaStream_t s) {
				cudaStream_t t = s;
				if (side) { cudaEventRecord(ev_fork[par][0], s); cudaStreamWaitEvent(side_stream, ev_fork[par][0], 0); t = side_stream; }
				HaloParams h{sV[D], peer_lo_Vs[D], (long long)((int)zb - z0) * plane, peer_lo_ghostE_off, comp, peer_lo_comp, plane,
				             d_halo_cnt, peer_lo_flagE, d_numTS, 1u};
				launch_k(k_halo_push, 296, 256, 0, t, h);
			});
		}
		// ---- H cells that depend on E values changed by the hooks
		if (fix_count) {
			lab("fix_H");
			L.push_back([this, par, i16](cudaStream_t s) {
				if (i16) launch1d(k_fix_H<uint16_t>, pFix[par], pFix[par].count, s);
				else launch1d(k_fix_H<uint32_t>, pFix[par], pFix[par].count, s);
			});
		}
		// ---- slab top plane: needs the neighbour's E plane (side stream; after fix_H, which may touch top-plane cells)
		if (multi && peer_hi) {
			lab("halo_wait_E");
			L.push_back([this, par, side](cudaStream_t s) {
				if (side) { cudaEventRecord(ev_fork[par][1], s); cudaStreamWaitEvent(side_stream, ev_fork[par][1], 0); s = side_stream; }
				WaitParams w{d_flagE, d_numTS, 1u, d_halo_err, halo_timeout_cycles()};
				launch_k(k_halo_wait, 1, 1, 0, s, w);
			});
			lab("update_H_top");
			L.push_back([this, par, i16, side](cudaStream_t s) {
				if (side) s = side_stream;
				const StencilParams& q = pHtop[par];
				const dim3 block(32, tune_rows);
				// one plane: the one-cell-per-thread kernel, out of place (source set -> destination set)
				const dim3 g((unsigned)((gn[0] - 1 + 31) / 32), (unsigned)((q.ny - 1 + tune_rows - 1) / tune_rows), 1);
				if (i16) { if (has_pml) launch_k(k_small_H<uint16_t, true>, g, block, 0, s, q); else launch_k(k_small_H<uint16_t, false>, g, block, 0, s, q); }
				else { if (has_pml) launch_k(k_small_H<uint32_t, true>, g, block, 0, s, q); else launch_k(k_small_H<uint32_t, false>, g, block, 0, s, q); }
			});
			if (lor_fused)
				for (size_t o = 0; o < lor_dev.size(); ++o) {
					if (!lor_dev[o].i_on || lor_dev[o].top_first >= lor_dev[o].i.count) continue;
					lab("lorentz_apply_I_top");
					L.push_back([this, par, o, side](cudaStream_t s) {
						if (side) s = side_stream;
						// Apply2Current for the dispersive cells of the top plane: the tail of the (z-sorted) list
						LorParams q = lor_dev[o].i;
						q.X = sI[par ^ 1]; q.first = lor_dev[o].top_first;
						launch1d(k_lorentz_apply, q, q.count - q.first, s);
					});
				}
		}
		// ---- H of the UPML shell, from the final E
		if (has_pml) {
			lab("shell_H");
			L.push_back([this, par, i16](cudaStream_t s) {
				const ShellParams& q = pShH[par];
				if (!q.nblocks) return;
				if (i16) launch_k(k_shell_H<uint16_t>, q.nblocks, dim3(32, 8), 0, s, q); else launch_k(k_shell_H<uint32_t>, q.nblocks, dim3(32, 8), 0, s, q);
			});
		}
		if (side && (peer_lo || peer_hi)) {
			lab("join_side");
			L.push_back([this, par](cudaStream_t s) { cudaEventRecord(ev_join[par], side_stream); cudaStreamWaitEvent(s, ev_join[par], 0); });
		}
		// ---- post-current hook of the TFSF box, then post / apply current hooks of the absorbing sheets: H of the
		//      destination set is final here
		if (pTfsf[1].groups) { lab("tfsf_I"); L.push_back([this, par](cudaStream_t s) { launch1d(k_tfsf, pTfsfD[par][1], pTfsfD[par][1].groups, s); }); }
		for (size_t a = sheet_dev.size(); a-- > 0;)
			if (sheet_dev[a].i.count) {
				lab("sheet_post_I");
				L.push_back([this, par, a](cudaStream_t s) { SheetParams q = pShI[par][a]; q.X = sI[par ^ 1]; launch1d(k_sheet_post, q, q.count, s); });
			}
		for (size_t a = sheet_dev.size(); a-- > 0;)
			if (sheet_dev[a].i.count) {
				lab("sheet_apply_I");
				L.push_back([this, par, a](cudaStream_t s) { SheetParams q = pShI[par][a]; q.X = sI[par ^ 1]; launch1d(k_sheet_apply_I, q, q.count, s); });
			}
		if (pExc[1].groups) { lab("excite_I"); L.push_back([this, par](cudaStream_t s) { launch1d(k_excite, pExcD[par][1], pExcD[par][1].groups, s); }); }
		if (multi && peer_hi) {
			lab("halo_push_H");
			L.push_back([this, D](cudaStream_t s) {
				HaloParams h{sI[D], peer_hi_Is[D], (long long)((int)ze - 1 - z0) * plane, peer_hi_ghostH_off, comp, peer_hi_comp, plane,
				             d_halo_cnt + 1, peer_hi_flagH, d_numTS, 1u};
				launch_k(k_halo_push, 296, 256, 0, s, h);
			});
		}
		lab("tick");
		L.push_back([this](cudaStream_t s) { launch_k(k_tick, 1, 1, 0, s, d_numTS); });
	}
	kernels_per_step = (unsigned)stepf[0].size();
}

// the voltage flux of the x-slab boxes sits in set 1 after an odd number of one-pass timesteps
void Engine::flux_sets_sync(bool to_set0)
{
	if (!d_flux_v2 || !(numTS_host & 1u)) return;
	for (int g = 0; g < 2; ++g) {
		if (xs_box[g] < 0) continue;
		const PmlBox& B = pE.box[xs_box[g]];
		const size_t n = (size_t)3 * B.n[0] * B.n[1] * B.n[2] * sizeof(float);
		if (to_set0) cudaMemcpyAsync(d_flux_v + B.off, d_flux_v2 + B.off, n, cudaMemcpyDeviceToDevice, stream);
		else cudaMemcpyAsync(d_flux_v2 + B.off, d_flux_v + B.off, n, cudaMemcpyDeviceToDevice, stream);
	}
	cudaStreamSynchronize(stream);
}

// schedule rebuild that keeps the state: the set of x-slab boxes may change with the options
// option "halo_timeout_s" (default 600 s): how long a slab waits for its neighbour's halo before giving up.
// Host-side skew between ranks (per-rank dump processing between bursts, uneven engine creation, a rank's launch
// queue filling before the other rank is enqueued) must stay below it.
long long Engine::halo_timeout_cycles() const
{
	int khz = 0;
	if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) != cudaSuccess || khz <= 0) khz = 2000000;
	return (long long)halo_timeout_s * (long long)khz * 1000ll;
}

int Engine::rebuild_schedule()
{
	CK(cudaSetDevice(device));
	CK(cudaStreamSynchronize(stream));
	if (fused_active) flux_sets_sync(true);
	build_schedule();
	if (fused_active) flux_sets_sync(false);
	return 0;
}

bool Engine::fused_auto_choice() const
{
	// measured crossover (profiles/experiments_r02.md #10): the one-cell-per-thread two-pass kernels equal the one-pass
	// schedule at 256^3 (16.8 M cells) and lose 8 % at 320^3
	return (long long)gn[0] * gn[1] * (long long)(ze - zb) >= fused_min_cells;
}

// switching between the one-pass and the two-pass schedule keeps the current fields: the two-pass
// kernels work in place on set 0
int Engine::set_fused_active(int req)
{
	const bool on = fused_possible && !edge_dirty && (req == 1 || (req < 0 && fused_auto_choice())) && !(lor_fused && !tma_req);
	CK(cudaSetDevice(device));
	CK(cudaStreamSynchronize(stream));
	if (fused_active) flux_sets_sync(true);
	if (fused_active && !on && (numTS_host & 1u)) {
		const size_t nfield = (size_t)3 * comp;
		CK(cudaMemcpyAsync(sV[0], sV[1], nfield * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		CK(cudaMemcpyAsync(sI[0], sI[1], nfield * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		CK(cudaStreamSynchronize(stream));
	}
	if (!fused_active && on && fused_possible && (numTS_host & 1u)) {
		// odd timestep count: the current state must sit in set 1
		const size_t nfield = (size_t)3 * comp;
		CK(cudaMemcpyAsync(sV[1], sV[0], nfield * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		CK(cudaMemcpyAsync(sI[1], sI[0], nfield * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		CK(cudaStreamSynchronize(stream));
	}
	fused_req = req;
	build_schedule();
	if (fused_active) flux_sets_sync(false);
	return 0;
}

void Engine::launch_probes(double* dst)
{
	if (!n_values) return;
	ProbeParams p = pProbe;
	p.V = sV[cur()]; p.I = sI[cur()];
	p.out = dst;
	const unsigned wpb = 4;
	launch_k(k_probes, (p.nprobes + wpb - 1) / wpb, wpb * 32, 0, stream, p);
	++kernels_launched;
}

int Engine::iterate(unsigned n)
{
	if (!finalized) return fail("iterate: engine not finalized");
	CK(cudaSetDevice(device));
	if (!probes_built && !h_probes.empty())
		if (build_probes()) return 1;
	if (ghost_open && n && release_ghosts()) return 1;
	if (fused_active && !sched_error.empty()) return fail(sched_error);
	for (unsigned it = 0; it < n; ++it) {
		if (fused_active) {
			const int par = (int)(numTS_host & 1u);
			if (use_graph) CK(cudaGraphLaunch(graphf_exec[par], stream));
			else for (auto& f : stepf[par]) f(stream);
		} else if (use_graph) {
			CK(cudaGraphLaunch(graph_exec, stream));
		} else {
			for (auto& f : step) f(stream);
		}
		kernels_launched += kernels_per_step;
		++numTS_host;
		if (rec_interval && numTS_host % rec_interval == 0 && rec_count < rec_cap && n_values) {
			launch_probes(d_series + (size_t)rec_count * n_values);
			rec_ts.push_back(numTS_host);
			++rec_count;
		}
	}
	CK(cudaGetLastError());
	return 0;
}

int Engine::set_option(const char* key, long long value)
{
	const std::string k = key ? key : "";
	if (finalized) CK(cudaSetDevice(device));
	if (k == "fused") {
		// 0: two-pass, 1: one-pass (if the hook set allows it), -1: automatic
		if (!finalized) { fused_req = value < 0 ? -1 : (value != 0); return 0; }
		if (value > 0 && edge_dirty) return 0;
		return set_fused_active(value < 0 ? -1 : (value != 0));
	}
	if (k == "halo_timeout_s") {
		halo_timeout_s = (int)std::max<long long>(1, std::min<long long>(86400, value));
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "small") { // two-pass stencil kernels with one cell per thread: 1 / 0 / -1 = automatic (small meshes)
		small_req = value < 0 ? -1 : (value != 0);
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "fused_min_cells") { // automatic schedule choice: one-pass from this many cells per GPU
		fused_min_cells = value;
		if (finalized && fused_req < 0) return set_fused_active(-1);
		return 0;
	}
	if (k == "small_max_cells") {
		small_max_cells = value;
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "overlap_halo") { // 1 (default): halo push / top plane of a z-slab on a side stream next to the UPML shell
		overlap_halo = value != 0;
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "pdl") { // 1: programmatic dependent launch between the kernels of a timestep (process-wide)
		g_pdl = value != 0;
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "skip_shell") { // 1 (default): the one-pass kernel skips UPML boxes that span whole planes / rows, 0: passes them through
		skip_req = value != 0;
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "xslab_zchunk") {
		xt_zchunk = (int)std::max<long long>(4, std::min<long long>(63, value)); // one shell bit per plane of a march
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "shell_zchunk") { // planes a UPML shell block marches (tuning aid)
		shell_zchunk = (int)std::max<long long>(1, std::min<long long>(64, value));
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "tma") {
		// 1: the one-pass kernel stages its inputs through TMA (default), 0: register-staged loads
		tma_req = value != 0;
		if (finalized && lor_fused) return set_fused_active(fused_req); // the register-staged kernel has no ADE instance: two-pass
		if (finalized) return rebuild_schedule();
		return 0;
	}
	if (k == "xslab") {
		// 1: thin UPML boxes at the x ends are updated by their own one-pass kernel k_xslab_EH, 0: shell launches
		// (default: measured faster, profiles/experiments_r01.md #12, #15)
		xslab_req = value <= 0 ? 0 : (value >= 2 ? 2 : 1); // 1: k_xslab_EH (register-staged), 2: k_xslab_tma
		if (finalized) return rebuild_schedule();
		return 0;
	}
	// (An "L2-blocked" launch order -- E kernel on a few planes, then the
	// H kernel one plane behind so that it reads from L2 -- was measured in round 1 and was 1.7x
	// SLOWER at 1024^3: see profiles/experiments_r01.md.)
	return fail("set_option: unknown key " + k);
}

int Engine::get_option(const char* key, long long* value)
{
	const std::string k = key ? key : "";
	if (!value) return fail("get_option: null pointer");
	if (k == "fused") { *value = fused_active ? 1 : 0; return 0; }
	if (k == "tma") { *value = (fused_active && tma_active) ? 1 : 0; return 0; }
	if (k == "pdl") { *value = g_pdl; return 0; }
	if (k == "small") { *value = !fused_active && small_active; return 0; }
	if (k == "skip_shell") { *value = fused_active ? skip_active : 0; return 0; }
	// rows / local planes the one-pass kernel works on (the all-UPML planes / rows at the mesh ends are skipped)
	if (k == "onepass_rows") { *value = fused_active ? pF[0].je - pF[0].jb : 0; return 0; }
	if (k == "onepass_planes") { *value = fused_active ? pF[0].kE1 - pF[0].kE0 : 0; return 0; }
	if (k == "xslab") { *value = fused_active ? (xs_box[0] >= 0) + (xs_box[1] >= 0) : 0; return 0; }
	if (k == "h2d_bytes") { *value = (long long)h2d_bytes; return 0; } // bytes copied host -> device since oems_cuda_create (counted at the copy calls)
	return fail("get_option: unknown key " + k);
}

// iterate(n) bracketed by CUDA events on the engine's stream; returns the device time in ms
int Engine::iterate_timed(unsigned n, double* ms)
{
	if (!finalized) return fail("iterate_timed: engine not finalized");
	CK(cudaSetDevice(device));
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a));
	CK(cudaEventCreate(&b));
	CK(cudaStreamSynchronize(stream));
	CK(cudaEventRecord(a, stream));
	int rc = iterate(n);
	if (rc) return rc;
	CK(cudaEventRecord(b, stream));
	CK(cudaEventSynchronize(b));
	float t = 0;
	CK(cudaEventElapsedTime(&t, a, b));
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	if (ms) *ms = t;
	return sync();
}

int Engine::sync()
{
	CK(cudaSetDevice(device));
	if (stream) {
		cudaError_t e = cudaStreamSynchronize(stream);
		if (e != cudaSuccess && peers_linked)
			return fail(std::string("device fault while stepping a z-slab (") + cudaGetErrorString(e) +
			            "): a halo wait gave up after option halo_timeout_s seconds without the neighbour's plane, or a kernel faulted");
		CK(e);
	}
	if (d_halo_err) {
		unsigned e = 0;
		CK(cudaMemcpy(&e, d_halo_err, sizeof(e), cudaMemcpyDeviceToHost));
		if (e) return fail("halo exchange timed out waiting for a neighbour GPU");
	}
	return 0;
}

int Engine::reset()
{
	if (!finalized) return fail("reset: engine not finalized");
	CK(cudaSetDevice(device));
	CK(cudaStreamSynchronize(stream));
	const size_t nfield = (size_t)3 * comp;
	for (int q = 0; q < 2; ++q) {
		if (!sV[q]) continue;
		CK(cudaMemsetAsync(sV[q], 0, nfield * sizeof(float), stream));
		CK(cudaMemsetAsync(sI[q], 0, nfield * sizeof(float), stream));
	}
	if (has_pml) {
		CK(cudaMemsetAsync(d_flux_v, 0, (size_t)flux_floats * sizeof(float), stream));
		CK(cudaMemsetAsync(d_flux_i, 0, (size_t)flux_floats * sizeof(float), stream));
		if (d_flux_v2) CK(cudaMemsetAsync(d_flux_v2, 0, (size_t)flux_floats * sizeof(float), stream));
	}
	for (SheetDev& D : sheet_dev) {
		CK(cudaMemsetAsync(D.v.store, 0, (size_t)D.v.count * sizeof(float), stream));
		if (D.i.count) CK(cudaMemsetAsync(D.i.store, 0, (size_t)D.i.count * sizeof(float), stream));
	}
	if (pMur.nplanes) {
		CK(cudaMemsetAsync(pMur.vP, 0, (size_t)pMur.total * sizeof(float), stream));
		CK(cudaMemsetAsync(pMur.vPP, 0, (size_t)pMur.total * sizeof(float), stream));
	}
	for (auto& D : lor_dev) {
		for (LorParams* P : {&D.v, &D.i}) {
			if (P->ade) CK(cudaMemsetAsync(P->ade, 0, (size_t)3 * P->count * sizeof(float), stream));
			if (P->lor_ade) CK(cudaMemsetAsync(P->lor_ade, 0, (size_t)3 * P->count * sizeof(float), stream));
		}
	}
	for (auto& R : rlc_dev) {
		CK(cudaMemsetAsync(R.Vd, 0, (size_t)3 * R.count * sizeof(float), stream));
		CK(cudaMemsetAsync(R.J, 0, (size_t)3 * R.count * sizeof(float), stream));
		CK(cudaMemsetAsync(R.Il, 0, (size_t)R.count * sizeof(float), stream));
	}
	if (ss_on) {
		CK(cudaMemsetAsync(pSs.rec, 0, (size_t)2 * ss_period * pSs.count * sizeof(double), stream));
		CK(cudaMemsetAsync(pSs.snap, 0, (size_t)2 * ss_period * pSs.count * sizeof(double), stream));
		CK(cudaMemsetAsync(pSs.energy, 0, 4 * sizeof(double), stream));
		CK(cudaMemsetAsync(pSs.info, 0, 2 * sizeof(unsigned), stream));
	}
	CK(cudaMemsetAsync(d_numTS, 0, sizeof(unsigned), stream));
	CK(cudaMemsetAsync(d_flagE, 0, FLAG_WORDS * sizeof(unsigned), stream));
	ghost_seq = 0; ghost_ts = 0; ghost_open = false;
	numTS_host = 0;
	rec_count = 0;
	rec_ts.clear();
	CK(cudaStreamSynchronize(stream));
	return 0;
}

// ------------------------------------------------------------------------------ probes
int Engine::add_probe_voltage(const unsigned start[3], const unsigned stop[3], int* id)
{
	if (rec_count) return fail("probes cannot be added while a recorded series holds samples");
	probes_built = false;
	for (int a = 0; a < 3; ++a)
		if (start[a] >= gn[a] || stop[a] >= gn[a]) return fail("add_probe_voltage: position outside the mesh");
	ProbeHost P;
	P.kind = 0;
	// engine_interface_fdtd.cpp:206-232
	if (((start[0] != stop[0]) + (start[1] != stop[1]) + (start[2] != stop[2])) == 1) {
		for (int n = 0; n < 3; ++n) {
			if (start[n] < stop[n]) {
				unsigned pos[3] = {start[0], start[1], start[2]};
				for (; pos[n] < stop[n]; ++pos[n])
					if (owned(pos[2])) { P.off.push_back(n * comp + cell_off(pos[0], pos[1], pos[2])); P.sign.push_back(1); }
			} else {
				unsigned pos[3] = {stop[0], stop[1], stop[2]};
				for (; pos[n] < start[n]; ++pos[n])
					if (owned(pos[2])) { P.off.push_back(n * comp + cell_off(pos[0], pos[1], pos[2])); P.sign.push_back(-1); }
			}
		}
	}
	if (id) *id = (int)h_probes.size();
	n_values += P.kind >= 2 ? 3 : 1; // num_probe_values() is the slot of the next probe even before the device lists are rebuilt
	h_probes.push_back(std::move(P));
	return 0;
}

int Engine::add_probe_current(const unsigned start[3], const unsigned stop[3], int nd, const int si[3], const int ei[3], int* id)
{
	if (rec_count) return fail("probes cannot be added while a recorded series holds samples");
	probes_built = false;
	for (int a = 0; a < 3; ++a)
		if (start[a] >= gn[a] || stop[a] >= gn[a] || start[a] > stop[a]) return fail("add_probe_current: bad box");
	if (nd < 0 || nd > 2) return fail("add_probe_current: bad normal direction");
	ProbeHost P;
	P.kind = 1;
	auto term = [&](int n, unsigned x, unsigned y, unsigned z, int sgn) {
		if (owned(z)) { P.off.push_back(n * comp + cell_off(x, y, z)); P.sign.push_back((signed char)sgn); }
	};
	// Common/processcurrent.cpp:96-171
	switch (nd) {
	case 0:
		if (ei[0] && si[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) term(1, stop[0], i, start[2], 1);
		if (ei[0] && ei[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) term(2, stop[0], stop[1], i, 1);
		if (si[0] && ei[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) term(1, start[0], i, stop[2], -1);
		if (si[0] && si[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) term(2, start[0], start[1], i, -1);
		break;
	case 1:
		if (si[0] && si[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) term(2, start[0], start[1], i, 1);
		if (ei[1] && ei[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) term(0, i, stop[1], stop[2], 1);
		if (ei[0] && ei[1]) for (unsigned i = start[2] + 1; i <= stop[2]; ++i) term(2, stop[0], stop[1], i, -1);
		if (si[1] && si[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) term(0, i, start[1], start[2], -1);
		break;
	default:
		if (si[1] && si[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) term(0, i, start[1], start[2], 1);
		if (ei[0] && si[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) term(1, stop[0], i, start[2], 1);
		if (ei[1] && ei[2]) for (unsigned i = start[0] + 1; i <= stop[0]; ++i) term(0, i, stop[1], stop[2], -1);
		if (si[0] && ei[2]) for (unsigned i = start[1] + 1; i <= stop[1]; ++i) term(1, start[0], i, stop[2], -1);
		break;
	}
	if (id) *id = (int)h_probes.size();
	n_values += P.kind >= 2 ? 3 : 1; // num_probe_values() is the slot of the next probe even before the device lists are rebuilt
	h_probes.push_back(std::move(P));
	return 0;
}

int Engine::add_probe_field(int is_H, const unsigned pos[3], int* id)
{
	if (rec_count) return fail("probes cannot be added while a recorded series holds samples");
	probes_built = false;
	for (int a = 0; a < 3; ++a)
		if (pos[a] >= gn[a]) return fail("add_probe_field: position outside the mesh");
	ProbeHost P;
	P.kind = is_H ? 3 : 2;
	if (owned(pos[2]))
		for (int n = 0; n < 3; ++n) { P.off.push_back(n * comp + cell_off(pos[0], pos[1], pos[2])); P.sign.push_back(1); }
	if (id) *id = (int)h_probes.size();
	n_values += P.kind >= 2 ? 3 : 1; // num_probe_values() is the slot of the next probe even before the device lists are rebuilt
	h_probes.push_back(std::move(P));
	return 0;
}

int Engine::build_probes()
{
	if (!finalized) return fail("probes need a finalized engine");
	std::vector<long long> off;
	std::vector<signed char> sign;
	std::vector<unsigned> pstart, pvalue;
	std::vector<unsigned char> kind;
	n_values = 0;
	for (const ProbeHost& P : h_probes) {
		pstart.push_back((unsigned)off.size());
		pvalue.push_back(n_values);
		kind.push_back((unsigned char)P.kind);
		off.insert(off.end(), P.off.begin(), P.off.end());
		sign.insert(sign.end(), P.sign.begin(), P.sign.end());
		n_values += P.kind >= 2 ? 3 : 1;
	}
	pstart.push_back((unsigned)off.size());
	memset(&pProbe, 0, sizeof(pProbe));
	pProbe.V = d_V; pProbe.I = d_I;
	pProbe.term_off = upload(off); pProbe.term_sign = upload(sign);
	pProbe.pstart = upload(pstart); pProbe.pvalue = upload(pvalue); pProbe.pkind = upload(kind);
	pProbe.nprobes = (unsigned)h_probes.size();
	d_probe_now = dalloc<double>(n_values);
	if (rec_interval && rec_cap) {
		d_series = dalloc<double>((size_t)rec_cap * n_values);
		if (!d_series) return fail("out of device memory (probe series)");
	}
	probes_built = true;
	return 0;
}

int Engine::read_probes(double* out)
{
	if (!finalized) return fail("read_probes: engine not finalized");
	CK(cudaSetDevice(device));
	if (!probes_built && build_probes()) return 1;
	if (!n_values) return 0;
	launch_probes(d_probe_now);
	CK(cudaMemcpyAsync(out, d_probe_now, n_values * sizeof(double), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	return 0;
}

int Engine::record_probes(unsigned interval, unsigned max_samples)
{
	if (probes_built && max_samples > rec_cap && interval) {
		CK(cudaSetDevice(device));
		d_series = dalloc<double>((size_t)max_samples * n_values);
		if (!d_series) return fail("out of device memory (probe series)");
	}
	rec_interval = interval;
	rec_cap = interval ? max_samples : 0;
	rec_count = 0;
	rec_ts.clear();
	return 0;
}

int Engine::read_probe_series(double* out, unsigned* ts_out, unsigned cap, unsigned* n)
{
	CK(cudaSetDevice(device));
	const unsigned cnt = std::min(cap, rec_count);
	if (cnt && n_values) {
		CK(cudaMemcpyAsync(out, d_series, (size_t)cnt * n_values * sizeof(double), cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
	}
	if (ts_out) for (unsigned i = 0; i < cnt; ++i) ts_out[i] = rec_ts[i];
	if (n) *n = cnt;
	return 0;
}

int Engine::energy(double* e)
{
	if (!finalized) return fail("energy: engine not finalized");
	CK(cudaSetDevice(device));
	EnergyParams p;
	p.V = sV[cur()]; p.I = sI[cur()];
	p.nx = (int)gn[0]; p.ny = (int)gn[1];
	p.k0 = (int)zb - z0; p.k1 = (int)std::min(ze, gn[2] - 1) - z0;
	p.pitch = pitch; p.plane = plane; p.comp = comp;
	p.acc = d_energy;
	CK(cudaMemsetAsync(d_energy, 0, 2 * sizeof(double), stream));
	const long long rows = (long long)(p.k1 - p.k0) * (p.ny - 1);
	if (rows > 0) {
		const unsigned blocks = (unsigned)std::min<long long>((rows + 7) / 8, 148 * 8);
		launch_k(k_energy, blocks, dim3(32, 8), 0, stream, p);
		++kernels_launched;
	}
	double acc[2];
	CK(cudaMemcpyAsync(acc, d_energy, sizeof(acc), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	*e = 8.85418781762e-12 * acc[0] + 1.256637062e-6 * acc[1];
	return 0;
}

// ------------------------------------------------------------------------------ measurement aids
int Engine::fill_fields(unsigned long long seed)
{
	if (!finalized) return fail("fill_fields: engine not finalized");
	CK(cudaSetDevice(device));
	FillParams p;
	p.V = sV[cur()]; p.I = sI[cur()];
	p.nx = (int)gn[0]; p.ny = (int)gn[1]; p.nzg = (int)gn[2];
	p.z0 = z0; p.nzl = nzl;
	p.pitch = pitch; p.plane = plane; p.comp = comp;
	p.seed = seed;
	const long long rows = (long long)nzl * gn[1];
	const unsigned blocks = (unsigned)std::min<long long>((rows + 7) / 8, 148 * 16);
	launch_k(k_fill, blocks, dim3(32, 8), 0, stream, p);
	++kernels_launched;
	CK(cudaStreamSynchronize(stream));
	return 0;
}

int Engine::field_digest(int is_curr, unsigned long long* out)
{
	if (!finalized) return fail("field_digest: engine not finalized");
	if (!out) return fail("field_digest: null pointer");
	CK(cudaSetDevice(device));
	unsigned long long* d_acc = (unsigned long long*)d_energy; // 16 bytes of scratch
	DigestParams p;
	p.X = is_curr ? sI[cur()] : sV[cur()];
	p.nx = (int)gn[0]; p.ny = (int)gn[1];
	p.z0 = z0; p.k0 = (int)zb - z0; p.k1 = (int)ze - z0;
	p.pitch = pitch; p.plane = plane; p.comp = comp;
	p.acc = d_acc;
	CK(cudaMemsetAsync(d_acc, 0, sizeof(unsigned long long), stream));
	const long long rows = (long long)(p.k1 - p.k0) * gn[1];
	const unsigned blocks = (unsigned)std::min<long long>((rows + 7) / 8, 148 * 16);
	launch_k(k_digest, blocks, dim3(32, 8), 0, stream, p);
	++kernels_launched;
	CK(cudaMemcpyAsync(out, d_acc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	return 0;
}

// ------------------------------------------------------------------------------ dumps
int Engine::add_dump(int is_H, int interp, unsigned nx, unsigned ny, unsigned nz, const unsigned* px, const unsigned* py,
                     const unsigned* pz, const double* const el[3], const double* const del[3], int* id)
{
	if (!finalized) return fail("add_dump: engine not finalized");
	if (interp < 0 || interp > 2) return fail("add_dump: bad interpolation type");
	CK(cudaSetDevice(device));
	for (unsigned i = 0; i < nx; ++i) if (px[i] >= gn[0]) return fail("add_dump: x index outside the mesh");
	for (unsigned i = 0; i < ny; ++i) if (py[i] >= gn[1]) return fail("add_dump: y index outside the mesh");
	for (unsigned i = 0; i < nz; ++i) if (pz[i] >= gn[2]) return fail("add_dump: z index outside the mesh");
	DumpHost D;
	memset(&D.p, 0, sizeof(D.p));
	// a z-slab engine evaluates the z lines it owns: a contiguous piece [z_first, z_first + nz) of the caller's list
	// (oems_cuda_dump_own_range), in the caller's output order; the host concatenates the slabs along z
	D.z_first = 0;
	if (slab_set) {
		for (unsigned i = 1; i < nz; ++i) if (pz[i] <= pz[i - 1]) return fail("add_dump: on a z-slab engine the z lines must ascend");
		unsigned a = 0, b = nz;
		while (a < nz && pz[a] < zb) ++a;
		b = a;
		while (b < nz && pz[b] < ze) ++b;
		D.z_first = a; pz += a; nz = b - a;
	}
	D.count = (size_t)nx * ny * nz;
	D.p.V = d_V; D.p.I = d_I;
	D.p.px = upload(std::vector<unsigned>(px, px + nx));
	D.p.py = upload(std::vector<unsigned>(py, py + ny));
	D.p.pz = upload(std::vector<unsigned>(pz, pz + nz));
	for (int a = 0; a < 3; ++a) {
		D.p.el[a] = upload(std::vector<double>(el[a], el[a] + gn[a]));
		D.p.del[a] = upload(std::vector<double>(del[a], del[a] + gn[a]));
	}
	D.d_out = dalloc<float>(std::max<size_t>(1, 3 * D.count));
	D.p.out = D.d_out;
	D.p.is_H = is_H; D.p.interp = interp;
	D.p.onx = nx; D.p.ony = ny; D.p.onz = nz;
	D.p.nx = (int)gn[0]; D.p.ny = (int)gn[1]; D.p.gnz = (int)gn[2]; D.p.z0 = z0;
	D.p.pitch = pitch; D.p.plane = plane; D.p.comp = comp;
	D.h_pinned = nullptr;
	CK(cudaMallocHost(&D.h_pinned, std::max<size_t>(1, 3 * D.count) * sizeof(float)));
	if (id) *id = (int)dumps.size();
	dumps.push_back(D);
	return 0;
}

int Engine::dump_own_range(int id, unsigned* first, unsigned* n)
{
	if (id < 0 || id >= (int)dumps.size()) return fail("dump_own_range: bad id");
	if (first) *first = dumps[id].z_first;
	if (n) *n = dumps[id].p.onz;
	return 0;
}

int Engine::read_dump(int id, float* out)
{
	if (id < 0 || id >= (int)dumps.size()) return fail("read_dump: bad id");
	long long t = 0;
	DumpHost& D = dumps[id];
	if (read_dump_async(id, D.h_pinned, &t)) return 1;
	if (wait_ticket(t)) return 1;
	memcpy(out, D.h_pinned, 3 * D.count * sizeof(float));
	return 0;
}

// ProcessFieldsTD::Process without stalling the time loop (north_star (4), Common/processfields_td.cpp:50-91):
// k_dump gathers/interpolates at the CURRENT timestep on the engine stream; the D2H copy into the caller's
// page-locked buffer runs on a second stream.  Stepping may continue immediately (the kernel has captured the
// fields of this timestep in d_out); the ticket is waited for only when the host needs the data.
int Engine::read_dump_async(int id, float* pinned_out, long long* ticket)
{
	if (id < 0 || id >= (int)dumps.size()) return fail("read_dump_async: bad id");
	if (!pinned_out || !ticket) return fail("read_dump_async: null pointer");
	CK(cudaSetDevice(device));
	DumpHost& D = dumps[id];
	if (!copy_stream) CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
	if (!D.ev_computed) {
		CK(cudaEventCreateWithFlags(&D.ev_computed, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&D.ev_copied, cudaEventDisableTiming));
	}
	if (D.copy_pending) CK(cudaStreamWaitEvent(stream, D.ev_copied, 0)); // d_out still being read by the last copy
	if (ghosts_for_readout()) return 1;
	D.p.V = sV[cur()]; D.p.I = sI[cur()];
	if (D.count) { launch1d(k_dump, D.p, (long long)D.count, stream); ++kernels_launched; }
	CK(cudaEventRecord(D.ev_computed, stream));
	CK(cudaStreamWaitEvent(copy_stream, D.ev_computed, 0));
	if (D.count) CK(cudaMemcpyAsync(pinned_out, D.d_out, 3 * D.count * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
	CK(cudaEventRecord(D.ev_copied, copy_stream));
	D.copy_pending = true;
	++D.seq;
	*ticket = ((long long)id << 32) | (long long)(D.seq & 0xffffffffull);
	return 0;
}

int Engine::wait_ticket(long long ticket)
{
	const int id = (int)(ticket >> 32);
	if (id < 0 || id >= (int)dumps.size()) return fail("wait: bad ticket");
	DumpHost& D = dumps[id];
	// an older ticket of the same box is covered too: copies of one box are ordered on the copy stream
	CK(cudaSetDevice(device));
	if (D.copy_pending) {
		CK(cudaEventSynchronize(D.ev_copied));
		D.copy_pending = false;
	}
	return 0;
}

// ------------------------------------------------------------------------------ FD dumps
// ProcessFieldsFD (Common/processfields_fd.cpp:40-107): one complex<float> accumulator per
// frequency and dump point, kept on the device; D2H once at the end (PostProcess)
int Engine::add_fd_dump(int dump_id, unsigned nfreq, int* id)
{
	if (dump_id < 0 || dump_id >= (int)dumps.size()) return fail("add_fd_dump: bad dump id");
	if (nfreq == 0) return fail("add_fd_dump: no frequencies"); // the reference disables such a dump (processfields_fd.cpp:44-49)
	CK(cudaSetDevice(device));
	FdHost F;
	F.dump = dump_id; F.nfreq = nfreq; F.samples = 0;
	const size_t n = 3 * dumps[dump_id].count;
	F.d_acc = dalloc<float2>(std::max<size_t>(1, n * nfreq));
	F.d_w = dalloc<float2>((size_t)FdHost::RING * nfreq);
	if (!F.d_acc || !F.d_w) return fail("out of device memory (FD dump)");
	F.h_w = nullptr;
	CK(cudaMallocHost((void**)&F.h_w, (size_t)FdHost::RING * nfreq * sizeof(float2)));
	for (int r = 0; r < FdHost::RING; ++r) CK(cudaEventCreateWithFlags(&F.ev[r], cudaEventDisableTiming));
	if (n) CK(cudaMemsetAsync(F.d_acc, 0, n * nfreq * sizeof(float2), stream));
	if (id) *id = (int)fds.size();
	fds.push_back(F);
	return 0;
}

// one sample of ProcessFieldsFD::Process; only enqueues (the weights travel through a small page-locked ring,
// so the caller's buffer is free on return and the time loop is not stalled)
int Engine::fd_accumulate(int fd_id, const float* w)
{
	if (fd_id < 0 || fd_id >= (int)fds.size()) return fail("fd_accumulate: bad id");
	if (!w) return fail("fd_accumulate: null pointer");
	CK(cudaSetDevice(device));
	FdHost& F = fds[fd_id];
	DumpHost& D = dumps[F.dump];
	const int slot = (int)(F.samples % FdHost::RING);
	if (F.samples >= (unsigned)FdHost::RING) CK(cudaEventSynchronize(F.ev[slot])); // the copy of RING samples ago has left this slot
	float2* hw = F.h_w + (size_t)slot * F.nfreq;
	float2* dw = F.d_w + (size_t)slot * F.nfreq;
	memcpy(hw, w, F.nfreq * sizeof(float2));
	if (D.copy_pending) CK(cudaStreamWaitEvent(stream, D.ev_copied, 0));
	if (ghosts_for_readout()) return 1;
	D.p.V = sV[cur()]; D.p.I = sI[cur()];
	if (D.count) launch1d(k_dump, D.p, (long long)D.count, stream);
	CK(cudaMemcpyAsync(dw, hw, F.nfreq * sizeof(float2), cudaMemcpyHostToDevice, stream));
	CK(cudaEventRecord(F.ev[slot], stream));
	FdParams q;
	q.td = D.d_out; q.acc = F.d_acc; q.w = dw; q.n = (long long)(3 * D.count); q.nfreq = F.nfreq;
	if (D.count) { launch1d(k_fd_accumulate, q, q.n, stream); kernels_launched += 2; }
	++F.samples;
	return 0;
}

int Engine::read_fd(int fd_id, float* out, unsigned* samples)
{
	if (fd_id < 0 || fd_id >= (int)fds.size()) return fail("read_fd: bad id");
	CK(cudaSetDevice(device));
	FdHost& F = fds[fd_id];
	const size_t n = 3 * dumps[F.dump].count * F.nfreq;
	if (out && n) CK(cudaMemcpyAsync(out, F.d_acc, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	if (samples) *samples = F.samples;
	return 0;
}

// ------------------------------------------------------------------------------ mode matching
// ProcessModeMatch (Common/processmodematch.cpp): start/stop are the box AFTER InitProcess has
// sorted it and pulled it off the boundaries (lines 86-104); dist0/dist1 are the normalised
// m_ModeDist arrays [nl0][nl1], area the GetNodeArea values of the same points
int Engine::add_mode_match(int is_H, int ny, const unsigned start[3], const unsigned stop[3], const double* dist0,
                           const double* dist1, const double* area, const double* const el[3], const double* const del[3], int* id)
{
	if (!finalized) return fail("add_mode_match: engine not finalized");
	if (ny < 0 || ny > 2 || !dist0 || !dist1 || !area) return fail("add_mode_match: bad arguments");
	for (int a = 0; a < 3; ++a)
		if (start[a] > stop[a] || stop[a] >= gn[a]) return fail("add_mode_match: box outside the mesh");
	if (start[ny] != stop[ny]) return fail("add_mode_match: the box is not a surface normal to ny");
	CK(cudaSetDevice(device));
	const int nP = (ny + 1) % 3, nPP = (ny + 2) % 3;
	ModeParams M;
	memset(&M, 0, sizeof(M));
	M.d.V = d_V; M.d.I = d_I;
	for (int a = 0; a < 3; ++a) {
		M.d.el[a] = upload(std::vector<double>(el[a], el[a] + gn[a]));
		M.d.del[a] = upload(std::vector<double>(del[a], del[a] + gn[a]));
	}
	M.d.is_H = is_H; M.d.interp = 1; // NODE_INTERPOLATE, processmodematch.cpp:82
	M.d.nx = (int)gn[0]; M.d.ny = (int)gn[1]; M.d.gnz = (int)gn[2]; M.d.z0 = z0;
	M.d.pitch = pitch; M.d.plane = plane; M.d.comp = comp;
	M.ny = ny; M.line = (int)start[ny];
	M.startP = (int)start[nP]; M.startPP = (int)start[nPP];
	M.nl0 = stop[nP] - start[nP] + 1; M.nl1 = stop[nPP] - start[nPP] + 1;
	const size_t n = (size_t)M.nl0 * M.nl1;
	M.dist0 = upload(std::vector<double>(dist0, dist0 + n));
	M.dist1 = upload(std::vector<double>(dist1, dist1 + n));
	M.area = upload(std::vector<double>(area, area + n));
	M.out = dalloc<double>(3);
	M.own_z0 = (int)zb; M.own_z1 = (int)ze; // a z-slab engine integrates over the planes it owns (read_mode_match_raw)
	if (!M.dist0 || !M.dist1 || !M.area || !M.out) return fail("out of device memory (mode match)");
	if (id) *id = (int)modes.size();
	modes.push_back(M);
	return 0;
}

int Engine::read_mode_match_raw(int id, double out[3])
{
	if (id < 0 || id >= (int)modes.size()) return fail("read_mode_match: bad id");
	CK(cudaSetDevice(device));
	ModeParams& M = modes[id];
	if (ghosts_for_readout()) return 1;
	M.d.V = sV[cur()]; M.d.I = sI[cur()];
	launch_k(k_mode_match, 1, 32, 0, stream, M);
	++kernels_launched;
	CK(cudaMemcpyAsync(out, M.out, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	return 0;
}

int Engine::read_mode_match(int id, double out[2])
{
	double r[3];
	if (read_mode_match_raw(id, r)) return 1;
	out[0] = r[0]; out[1] = r[1];
	return 0;
}

// ------------------------------------------------------------------------------ complete ghost planes
// z-slab engines: before an interpolating readout both ghost planes are completed (all components of E and H, see
// k_ghost_push).  Only enqueues.  One exchange serves every readout of the same timestep.  Protocol per slab:
//   push my lowest / highest owned plane -> publish exchange number q in the neighbours' flags -> wait for theirs
//   ... readout kernels ...
//   (release, enqueued by the next iterate / exchange) publish "done reading q" -> wait for the neighbours' "done":
//   only then may this slab's time loop write into the neighbours' ghost planes again.
// Why the pushes cannot disturb a neighbour that lags: a slab that has finished timestep n has consumed its
// neighbours' step-n halos, so both neighbours are past the phases of step n that read the ghost values the push
// rewrites, and those values are rewritten with the same bits.
int Engine::exchange_ghosts()
{
	if (!finalized) return fail("exchange_ghosts: engine not finalized");
	if (!peers_linked) return 0;
	CK(cudaSetDevice(device));
	if (ghost_open && ghost_ts == numTS_host) return 0;
	if (ghost_open && release_ghosts()) return 1;
	if (peer_lo && !peer_lo_Is[0]) return fail("exchange_ghosts: the lower neighbour's buffers are not mapped");
	++ghost_seq;
	const int c = cur();
	if (peer_lo) {
		GhostPushParams g{sV[c], sI[c], peer_lo_Vs[c], peer_lo_Is[c], (long long)((int)zb - z0) * plane, peer_lo_ghostE_off, comp, peer_lo_comp,
		                  plane, d_flagE + FLAG_GCNT, peer_lo_flags + FLAG_G_HI, ghost_seq};
		launch_k(k_ghost_push, 64, 256, 0, stream, g);
	}
	if (peer_hi) {
		GhostPushParams g{sV[c], sI[c], peer_hi_Vs[c], peer_hi_Is[c], (long long)((int)ze - 1 - z0) * plane, peer_hi_ghostH_off, comp, peer_hi_comp,
		                  plane, d_flagE + FLAG_GCNT + 1, peer_hi_flags + FLAG_G_LO, ghost_seq};
		launch_k(k_ghost_push, 64, 256, 0, stream, g);
	}
	FlagParams w{{peer_lo ? d_flagE + FLAG_G_LO : nullptr, peer_hi ? d_flagE + FLAG_G_HI : nullptr}, ghost_seq, d_halo_err, halo_timeout_cycles()};
	launch_k(k_flag_wait, 1, 1, 0, stream, w);
	kernels_launched += 1 + (peer_lo != nullptr) + (peer_hi != nullptr);
	ghost_open = true;
	ghost_ts = numTS_host;
	CK(cudaGetLastError());
	return 0;
}

int Engine::release_ghosts()
{
	if (!ghost_open) return 0;
	CK(cudaSetDevice(device));
	FlagParams a{{peer_lo ? peer_lo_flags + FLAG_ACK_HI : nullptr, peer_hi ? peer_hi_flags + FLAG_ACK_LO : nullptr}, ghost_seq, d_halo_err, 0};
	launch_k(k_flag_set, 1, 1, 0, stream, a);
	FlagParams w{{peer_lo ? d_flagE + FLAG_ACK_LO : nullptr, peer_hi ? d_flagE + FLAG_ACK_HI : nullptr}, ghost_seq, d_halo_err, halo_timeout_cycles()};
	launch_k(k_flag_wait, 1, 1, 0, stream, w);
	kernels_launched += 2;
	ghost_open = false;
	CK(cudaGetLastError());
	return 0;
}

// readers call this: a slab with neighbours needs complete ghost planes of the current timestep
int Engine::ghosts_for_readout()
{
	if (!peers_linked) return 0;
	return exchange_ghosts();
}

// ------------------------------------------------------------------------------ field access
int Engine::get_field(int is_curr, unsigned n, unsigned x, unsigned y, unsigned z, float* v)
{
	if (!finalized) return fail("get_field: engine not finalized");
	if (n > 2 || x >= gn[0] || y >= gn[1] || !held(z)) return fail("get_field: position not on this engine");
	CK(cudaSetDevice(device));
	const float* base = is_curr ? sI[cur()] : sV[cur()];
	CK(cudaMemcpyAsync(v, base + n * comp + cell_off(x, y, z), sizeof(float), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	return 0;
}
int Engine::set_field(int is_curr, unsigned n, unsigned x, unsigned y, unsigned z, float v)
{
	if (!finalized) return fail("set_field: engine not finalized");
	if (n > 2 || x >= gn[0] || y >= gn[1] || !held(z)) return fail("set_field: position not on this engine");
	CK(cudaSetDevice(device));
	float* base = is_curr ? sI[cur()] : sV[cur()];
	CK(cudaMemcpyAsync(base + n * comp + cell_off(x, y, z), &v, sizeof(float), cudaMemcpyHostToDevice, stream));
	CK(cudaStreamSynchronize(stream));
	if (is_curr && (x == gn[0] - 1 || y == gn[1] - 1 || z == gn[2] - 1)) mark_edge_dirty();
	return 0;
}

int Engine::get_fields(int is_curr, float* out)
{
	if (!finalized) return fail("get_fields: engine not finalized");
	CK(cudaSetDevice(device));
	std::vector<float> h((size_t)3 * comp);
	CK(cudaMemcpyAsync(h.data(), is_curr ? sI[cur()] : sV[cur()], h.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	const unsigned nx = gn[0], ny = gn[1];
#pragma omp parallel for collapse(2) schedule(static)
	for (int n = 0; n < 3; ++n)
		for (unsigned i = 0; i < nx; ++i)
			for (unsigned j = 0; j < ny; ++j)
				for (int k = 0; k < nzl; ++k)
					out[(((size_t)n * nx + i) * ny + j) * nzl + k] = h[n * comp + (long long)k * plane + (long long)j * pitch + i];
	return 0;
}
int Engine::set_fields(int is_curr, const float* in)
{
	if (!finalized) return fail("set_fields: engine not finalized");
	CK(cudaSetDevice(device));
	std::vector<float> h((size_t)3 * comp, 0.0f);
	const unsigned nx = gn[0], ny = gn[1];
#pragma omp parallel for collapse(2) schedule(static)
	for (int n = 0; n < 3; ++n)
		for (unsigned i = 0; i < nx; ++i)
			for (unsigned j = 0; j < ny; ++j)
				for (int k = 0; k < nzl; ++k)
					h[n * comp + (long long)k * plane + (long long)j * pitch + i] = in[(((size_t)n * nx + i) * ny + j) * nzl + k];
	CK(cudaMemcpyAsync(is_curr ? sI[cur()] : sV[cur()], h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
	CK(cudaStreamSynchronize(stream));
	if (is_curr) mark_edge_dirty();
	return 0;
}

// The currents on the last line of each direction are never written by the engine (engine.cpp:179-183)
// and start at +0; with a +0 flux the UPML pre/post pair maps +0 to +0, so k_upml_untouched_H is
// only scheduled once the caller has poked a current there (SetCurr / set_fields).
void Engine::mark_edge_dirty()
{
	if (edge_dirty || !edge_possible) return;
	if (!pEdge.count && build_edge_list()) return;
	if (!pEdge.count) return;
	edge_dirty = true;
	set_fused_active(0); // the one-pass schedule does not carry this rare path; also rebuilds the schedule
}

int Engine::get_upml_flux(int box, int is_curr, float* out)
{
	if (!finalized) return fail("get_upml_flux: engine not finalized");
	if (box < 0 || box >= (int)h_upml.size()) return fail("get_upml_flux: bad box");
	CK(cudaSetDevice(device));
	const UpmlBoxHost& B = h_upml[box];
	const size_t full = (size_t)3 * B.n[0] * B.n[1] * B.n[2];
	memset(out, 0, full * sizeof(float));
	if (B.ln[2] == 0) return 0;
	const long long cs = (long long)B.ln[0] * B.ln[1] * B.ln[2];
	std::vector<float> h((size_t)3 * cs);
	const float* src = is_curr ? d_flux_i : d_flux_v;
	if (!is_curr && fused_active && d_flux_v2 && (numTS_host & 1u))
		for (int g = 0; g < 2; ++g)
			if (xs_box[g] >= 0 && pE.box[xs_box[g]].off == B.flux_off) src = d_flux_v2; // x-slab box: the current flux sits in set 1
	CK(cudaMemcpyAsync(h.data(), src + B.flux_off, h.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	for (int n = 0; n < 3; ++n)
		for (int li = 0; li < B.ln[0]; ++li)
			for (int lj = 0; lj < B.ln[1]; ++lj)
				for (int lk = 0; lk < B.ln[2]; ++lk) {
					const unsigned gk = B.gz0 + lk - B.start[2];
					out[(((size_t)n * B.n[0] + li) * B.n[1] + lj) * B.n[2] + gk] = h[n * cs + ((long long)lk * B.ln[1] + lj) * B.ln[0] + li];
				}
	return 0;
}

int Engine::get_stats(oems_cuda_stats* s)
{
	memset(s, 0, sizeof(*s));
	s->n_unique = n_unique;
	s->index_bytes = index_bytes;
	s->hbm_bytes = hbm_bytes;
	s->kernels_launched = kernels_launched;
	s->kernels_per_step = kernels_per_step;
	s->pml_cells_lo = (unsigned)(pml_cells & 0xffffffffu);
	s->pml_cells_hi = (unsigned)(pml_cells >> 32);
	s->uses_graph = use_graph ? 1 : 0;
	return 0;
}

// planes marched per block: long chunks amortise the k-1 plane reload, short chunks give the
// block scheduler enough blocks to hide the tail on thin slabs (profiles/experiments_r01.md #5)
int Engine::auto_zchunk() const
{
	const long long planes = (long long)(ze - zb);
	const long long blocks_xy = (long long)((pitch / 4 + 31) / 32) * ((gn[1] + tune_rows - 1) / tune_rows);
	const long long want = 16LL * 1184; // ~16 waves of resident 128-thread blocks on 148 SMs
	for (int zc = 32; zc > 1; zc /= 2)
		if (blocks_xy * ((planes + zc - 1) / zc) >= want) return zc;
	return 1; // small meshes: one plane per block, maximum parallelism, shortest dependency chain
}

int Engine::set_tuning(int rows, int zchunk, int graph_on)
{
	if (rows > 0) {
		if (rows > 8) return fail("set_tuning: at most 8 rows per block (256 threads)");
		tune_rows = rows;
	}
	if (zchunk > 0) tune_zchunk = zchunk;
	if (graph_on >= 0) tune_graph = graph_on;
	if (finalized) {
		CK(cudaSetDevice(device));
		CK(cudaStreamSynchronize(stream));
		pE.zchunk = pH.zchunk = tune_zchunk > 0 ? tune_zchunk : auto_zchunk();
		build_schedule();
	}
	return 0;
}


// times every entry of the per-timestep schedule with CUDA events on the launch stream
// (bench.py roofline: the kernels run on this engine's own stream, which torch events do not see)
int Engine::time_schedule(unsigned n_ts, double* ms_out, unsigned cap, unsigned* n_entries)
{
	if (!finalized) return fail("time_schedule: engine not finalized");
	CK(cudaSetDevice(device));
	const size_t ne = fused_active ? stepf[0].size() : step.size();
	if (n_entries) *n_entries = (unsigned)ne;
	if (cap < ne) return fail("time_schedule: output too small");
	std::vector<cudaEvent_t> ev(ne + 1);
	for (auto& e : ev) CK(cudaEventCreate(&e));
	std::vector<double> acc(ne, 0.0);
	for (unsigned it = 0; it < n_ts; ++it) {
		for (size_t q = 0; q < ne; ++q) {
			CK(cudaEventRecord(ev[q], stream));
			if (fused_active) stepf[numTS_host & 1u][q](stream);
			else step[q](stream);
		}
		CK(cudaEventRecord(ev[ne], stream));
		CK(cudaStreamSynchronize(stream));
		for (size_t q = 0; q < ne; ++q) {
			float ms = 0;
			CK(cudaEventElapsedTime(&ms, ev[q], ev[q + 1]));
			acc[q] += ms;
		}
		kernels_launched += kernels_per_step;
		++numTS_host;
	}
	for (auto& e : ev) cudaEventDestroy(e);
	for (size_t q = 0; q < ne; ++q) ms_out[q] = n_ts ? acc[q] / n_ts : 0.0;
	return 0;
}

// ------------------------------------------------------------------------------ multi-GPU
struct IpcBlob {
	cudaIpcMemHandle_t hV, hI, hFlags, hV1, hI1;
	int has_set1;
	long long comp, plane;
	int z0, nzl, zb, ze, pitch, ny;
	int device;
};
static_assert(sizeof(IpcBlob) <= OEMS_IPC_BYTES, "IPC blob too large");

int Engine::export_ipc(unsigned char* out)
{
	if (!finalized) return fail("export_ipc: engine not finalized");
	CK(cudaSetDevice(device));
	IpcBlob b;
	memset(&b, 0, sizeof(b));
	CK(cudaIpcGetMemHandle(&b.hV, d_V));
	CK(cudaIpcGetMemHandle(&b.hI, d_I));
	CK(cudaIpcGetMemHandle(&b.hFlags, d_flagE));
	b.has_set1 = sV[1] != nullptr;
	if (b.has_set1) {
		CK(cudaIpcGetMemHandle(&b.hV1, sV[1]));
		CK(cudaIpcGetMemHandle(&b.hI1, sI[1]));
	}
	b.comp = comp; b.plane = plane; b.z0 = z0; b.nzl = nzl; b.zb = (int)zb; b.ze = (int)ze; b.pitch = pitch; b.ny = (int)gn[1];
	b.device = device;
	memset(out, 0, OEMS_IPC_BYTES);
	memcpy(out, &b, sizeof(b));
	return 0;
}

int Engine::open_peers(const unsigned char* lower, const unsigned char* upper)
{
	if (!finalized) return fail("open_peers: engine not finalized");
	CK(cudaSetDevice(device));
	CK(cudaStreamSynchronize(stream));
	peer_lo = peer_hi = nullptr;
	if (lower) {
		IpcBlob b;
		memcpy(&b, lower, sizeof(b));
		if (b.pitch != pitch || b.ny != (int)gn[1] || b.ze != (int)zb) return fail("open_peers: lower neighbour does not match");
		void *pV = nullptr, *pF = nullptr;
		CK(cudaIpcOpenMemHandle(&pV, b.hV, cudaIpcMemLazyEnablePeerAccess));
		CK(cudaIpcOpenMemHandle(&pF, b.hFlags, cudaIpcMemLazyEnablePeerAccess));
		ipc_opened.push_back(pV); ipc_opened.push_back(pF);
		peer_lo_V = (float*)pV;
		peer_lo_Vs[0] = peer_lo_V;
		if (b.has_set1) {
			void* pV1 = nullptr;
			CK(cudaIpcOpenMemHandle(&pV1, b.hV1, cudaIpcMemLazyEnablePeerAccess));
			ipc_opened.push_back(pV1);
			peer_lo_Vs[1] = (float*)pV1;
		} else fused_possible = false;
		peer_lo_flagE = (unsigned*)pF; // neighbour's flagE
		peer_lo_flags = (unsigned*)pF;
		{   // the other field of the neighbour: only the ghost exchange of the readout writes it
			void* pI = nullptr;
			CK(cudaIpcOpenMemHandle(&pI, b.hI, cudaIpcMemLazyEnablePeerAccess));
			ipc_opened.push_back(pI);
			peer_lo_Is[0] = (float*)pI; peer_lo_Is[1] = nullptr;
			if (b.has_set1) {
				void* pI1 = nullptr;
				CK(cudaIpcOpenMemHandle(&pI1, b.hI1, cudaIpcMemLazyEnablePeerAccess));
				ipc_opened.push_back(pI1);
				peer_lo_Is[1] = (float*)pI1;
			}
		}
		peer_lo_comp = b.comp;
		peer_lo_ghostE_off = (long long)((int)zb - b.z0) * plane;
		peer_lo = this; // marker: has a lower neighbour
	}
	if (upper) {
		IpcBlob b;
		memcpy(&b, upper, sizeof(b));
		if (b.pitch != pitch || b.ny != (int)gn[1] || b.zb != (int)ze) return fail("open_peers: upper neighbour does not match");
		void *pI = nullptr, *pF = nullptr;
		CK(cudaIpcOpenMemHandle(&pI, b.hI, cudaIpcMemLazyEnablePeerAccess));
		CK(cudaIpcOpenMemHandle(&pF, b.hFlags, cudaIpcMemLazyEnablePeerAccess));
		ipc_opened.push_back(pI); ipc_opened.push_back(pF);
		peer_hi_I = (float*)pI;
		peer_hi_Is[0] = peer_hi_I;
		if (b.has_set1) {
			void* pI1 = nullptr;
			CK(cudaIpcOpenMemHandle(&pI1, b.hI1, cudaIpcMemLazyEnablePeerAccess));
			ipc_opened.push_back(pI1);
			peer_hi_Is[1] = (float*)pI1;
		} else fused_possible = false;
		peer_hi_flagH = (unsigned*)pF + 1; // neighbour's flagH
		peer_hi_flags = (unsigned*)pF;
		{
			void* pV = nullptr;
			CK(cudaIpcOpenMemHandle(&pV, b.hV, cudaIpcMemLazyEnablePeerAccess));
			ipc_opened.push_back(pV);
			peer_hi_Vs[0] = (float*)pV; peer_hi_Vs[1] = nullptr;
			if (b.has_set1) {
				void* pV1 = nullptr;
				CK(cudaIpcOpenMemHandle(&pV1, b.hV1, cudaIpcMemLazyEnablePeerAccess));
				ipc_opened.push_back(pV1);
				peer_hi_Vs[1] = (float*)pV1;
			}
		}
		peer_hi_comp = b.comp;
		peer_hi_ghostH_off = (long long)((int)ze - 1 - b.z0) * plane;
		peer_hi = this;
	}
	peers_linked = lower || upper;
	build_schedule();
	return 0;
}

int Engine::link_peers(Engine* lower, Engine* upper)
{
	if (!finalized) return fail("link_peers: engine not finalized");
	CK(cudaSetDevice(device));
	CK(cudaStreamSynchronize(stream));
	peer_lo = lower; peer_hi = upper;
	for (Engine* o : {lower, upper}) {
		if (!o) continue;
		if (!o->finalized) return fail("link_peers: neighbour not finalized");
		if (o->device != device) {
			int can = 0;
			CK(cudaDeviceCanAccessPeer(&can, device, o->device));
			if (!can) return fail("link_peers: no peer access between the two GPUs");
			cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return check(e, "cudaDeviceEnablePeerAccess");
			cudaGetLastError();
		}
	}
	if (lower) {
		if (lower->pitch != pitch || lower->ze != zb) return fail("link_peers: lower neighbour does not match");
		peer_lo_V = lower->d_V; peer_lo_flagE = lower->d_flagE; peer_lo_comp = lower->comp;
		peer_lo_Vs[0] = lower->sV[0]; peer_lo_Vs[1] = lower->sV[1];
		peer_lo_Is[0] = lower->sI[0]; peer_lo_Is[1] = lower->sI[1];
		peer_lo_flags = lower->d_flagE;
		if (!lower->sV[1] || !lower->fused_possible) fused_possible = false;
		peer_lo_ghostE_off = (long long)((int)zb - lower->z0) * plane;
	}
	if (upper) {
		if (upper->pitch != pitch || upper->zb != ze) return fail("link_peers: upper neighbour does not match");
		peer_hi_I = upper->d_I; peer_hi_flagH = upper->d_flagH; peer_hi_comp = upper->comp;
		peer_hi_Is[0] = upper->sI[0]; peer_hi_Is[1] = upper->sI[1];
		peer_hi_Vs[0] = upper->sV[0]; peer_hi_Vs[1] = upper->sV[1];
		peer_hi_flags = upper->d_flagE;
		if (!upper->sI[1] || !upper->fused_possible) fused_possible = false;
		peer_hi_ghostH_off = (long long)((int)ze - 1 - upper->z0) * plane;
	}
	peers_linked = lower || upper;
	build_schedule();
	return 0;
}

#include "abi.inc"
