// engine.h -- host-side state of one B200 FDTD engine instance (one GPU, one z-slab).
// The C ABI in include/openems_b200.h is a thin wrapper over this class (abi.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/openems_b200.h"
#include "kernels.cuh"
#include "kernels_fused.cuh"
#include "kernels_fused_tma.cuh"
#include "kernels_xslab.cuh"
#include "kernels_xslab_tma.cuh"

struct UpmlBoxHost {
	unsigned start[3], n[3];          // global
	std::vector<float> c[6];          // dense aux coefficients (may be empty with a compressed operator)
	// part of the box held by this GPU (owned planes only)
	int ls[3], ln[3];                 // local start (x, y, LOCAL z) and lines; ln[2]==0: not on this GPU
	unsigned gz0;                     // global z of the first held plane of the box
	long long flux_off;
};
struct MurHost {
	int ny; unsigned line, shift, n[2], start_ts;
	std::vector<float> cP, cPP;
};
struct LorHost {
	unsigned count;
	std::vector<unsigned> pos;        // [3][count]
	std::vector<float> c[6];          // v_int v_ext v_lor i_int i_ext i_lor, [3][count] or empty
};
struct RlcHost {
	unsigned count;
	std::vector<int> dir;
	std::vector<unsigned> pos;
	std::vector<float> c[9];
};
struct ExcHost {
	std::vector<unsigned> idx[3], dir, delay;
	std::vector<float> amp;
};
struct ProbeHost {
	int kind; // 0 voltage, 1 current, 2 raw E, 3 raw H
	std::vector<long long> off;
	std::vector<signed char> sign;
};
struct DumpHost {
	DumpParams p;
	size_t count;
	unsigned z_first = 0;   // z-slab engines: first entry of the caller's z list this engine evaluates
	float* d_out;
	float* h_pinned;
	std::vector<void*> dev_allocs;
	// asynchronous read-out (oems_cuda_read_dump_async): k_dump runs on the engine stream, the D2H copy on the
	// copy stream; ev_computed orders the copy after the kernel, ev_copied is the ticket and also keeps the
	// next k_dump of this box from overwriting d_out before the copy has left
	cudaEvent_t ev_computed = nullptr, ev_copied = nullptr;
	unsigned long long seq = 0;
	bool copy_pending = false;
};

struct TfsfHost {     // Operator_Ext_TFSF tables, index (n*2+l)*2+c
	bool on = false;
	unsigned start[3], stop[3];
	int active[3][2];
	std::vector<unsigned> delay[2][12];
	std::vector<float> dd[2][12], amp[2][12];
};

struct SheetHost {    // one local absorbing sheet (Operator_Ext_Absorbing_BC)
	int ny, type, positive;
	unsigned x0[3], x1[3];
	std::vector<float> K1P, K1PP, K2P, K2PP;
};
struct SheetDev { SheetParams v, i; };

struct FdHost {
	static const int RING = 4;
	int dump;          // time-domain dump this spectrum is taken from
	unsigned nfreq;
	float2* d_acc;     // [nfreq][3*count]
	float2* d_w;       // [RING][nfreq]: weights of the last RING samples
	float2* h_w;       // the same in page-locked host memory (source of the asynchronous H2D copies)
	cudaEvent_t ev[RING]; // slot s may be rewritten once ev[s] has passed
	unsigned samples;
};

struct LorDev {
	LorParams v, i;
	bool v_on, i_on;
	int2* d_rows = nullptr;          // one-pass schedule: per (local plane, row) {x0 | n << 16, first list index}
	std::vector<int2> h_rows;
	unsigned top_first = 0;          // list entries [top_first, count) lie on the slab's top owned plane
};

class Engine {
public:
	Engine(unsigned nx, unsigned ny, unsigned nz, int device);
	~Engine();

	std::string err;
	int fail(const std::string& m) { err = m; return 1; }

	int set_slab(unsigned zb, unsigned ze);
	int set_operator_dense(const float* vv, const float* vi, const float* ii, const float* iv);
	int set_operator_planes(unsigned nu, const oems_coeff_entry* table, unsigned n_uplanes, const void* uplanes, const unsigned* plane_of_z, int ib);
	int set_operator_compressed(unsigned n_unique, const oems_coeff_entry* table, const void* index, int index_bytes);
	int set_signal(const float* sv, const float* si, unsigned len, unsigned period);
	int add_excitation(int is_curr, unsigned count, const unsigned* idx3, const unsigned* dir, const float* amp, const unsigned* delay);
	int add_upml(const unsigned start[3], const unsigned n[3], const float* const c[6]);
	int add_mur(int ny, unsigned line, unsigned shift, const unsigned n[2], const float* cP, const float* cPP, unsigned start_ts);
	int add_lorentz(unsigned count, const unsigned* pos3, const float* const c[6]);
	int add_rlc(unsigned count, const int* dir, const unsigned* pos3, const float* const c[9]);
	int add_steadystate(unsigned period_ts, unsigned count, const unsigned* pos3, const unsigned* dir);
	int steadystate_check(double* last_diff, unsigned* n_checks);
	int steadystate_raw(unsigned info[2], double en[4], double* snap, unsigned cap, unsigned* count);
	static int steadystate_eval(unsigned period, unsigned count, const unsigned info[2], const double en[4], const double* snap, double* last_diff);
	int dump_own_range(int id, unsigned* first, unsigned* n);
	int read_mode_match_raw(int id, double out[3]);
	int finalize();

	int iterate(unsigned n);
	int iterate_timed(unsigned n, double* ms);
	int sync();
	int reset();
	unsigned num_ts() const { return numTS_host; }

	int add_probe_voltage(const unsigned start[3], const unsigned stop[3], int* id);
	int add_probe_current(const unsigned start[3], const unsigned stop[3], int nd, const int si[3], const int ei[3], int* id);
	int add_probe_field(int is_H, const unsigned pos[3], int* id);
	unsigned num_probe_values() const { return n_values; }
	int read_probes(double* out);
	int record_probes(unsigned interval, unsigned max_samples);
	int read_probe_series(double* out, unsigned* ts_out, unsigned cap, unsigned* n);
	int energy(double* e);
	int fill_fields(unsigned long long seed);
	int field_digest(int is_curr, unsigned long long* out);
	int add_dump(int is_H, int interp, unsigned nx, unsigned ny, unsigned nz, const unsigned* px, const unsigned* py,
	             const unsigned* pz, const double* const el[3], const double* const del[3], int* id);
	int read_dump(int id, float* out);
	int read_dump_async(int id, float* pinned_out, long long* ticket);
	int wait_ticket(long long ticket);
	int add_fd_dump(int dump_id, unsigned nfreq, int* id);
	int fd_accumulate(int fd_id, const float* w);
	int read_fd(int fd_id, float* out, unsigned* samples);
	std::vector<FdHost> fds;
	int add_absorbing_sheet(int ny, const unsigned x0[3], const unsigned x1[3], int positive, int type, const float* K1P,
	                        const float* K1PP, const float* K2P, const float* K2PP);
	int build_sheets();
	int set_tfsf(const unsigned start[3], const unsigned stop[3], const int active[6], const unsigned* const* vdelay, const float* const* vdd,
	             const float* const* vamp, const unsigned* const* cdelay, const float* const* cdd, const float* const* camp);
	int build_tfsf();
	TfsfHost h_tfsf;
	TfsfParams pTfsf[2];          // voltage / current update lists
	TfsfParams pTfsfD[2][2];      // per parity of the one-pass schedule (destination set)
	std::vector<SheetHost> h_sheet;
	std::vector<SheetDev> sheet_dev;
	SheetParams pShV[2][8], pShI[2][8]; // per parity of the one-pass schedule: [0] works on the source set, [1] on the destination set
	int add_mode_match(int is_H, int ny, const unsigned start[3], const unsigned stop[3], const double* dist0, const double* dist1,
	                   const double* area, const double* const el[3], const double* const del[3], int* id);
	int read_mode_match(int id, double out[2]);
	std::vector<ModeParams> modes;
	int get_field(int is_curr, unsigned n, unsigned x, unsigned y, unsigned z, float* v);
	int set_field(int is_curr, unsigned n, unsigned x, unsigned y, unsigned z, float v);
	int get_fields(int is_curr, float* out);
	int set_fields(int is_curr, const float* in);
	int get_upml_flux(int box, int is_curr, float* out);
	int get_stats(oems_cuda_stats* s);
	int set_tuning(int rows, int zchunk, int use_graph);
	int time_schedule(unsigned n_ts, double* ms_out, unsigned cap, unsigned* n_entries);
	const char* schedule_label(unsigned i) const
	{
		const std::vector<std::string>& L = fused_active ? labelsf : labels;
		return i < L.size() ? L[i].c_str() : "";
	}
	int export_ipc(unsigned char* out);
	int open_peers(const unsigned char* lower, const unsigned char* upper);
	int link_peers(Engine* lower, Engine* upper);

private:
	// geometry
	unsigned gn[3];
	int device;
	unsigned zb, ze;      // owned global planes [zb, ze)
	int z0;               // global z of local plane 0
	int nzl;              // local planes (owned + ghosts)
	int pitch;
	long long plane, comp;
	bool finalized = false, slab_set = false;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr; // D2H of dumps, overlapping the time loop

	// host staging
	std::vector<float> h_dense[4];    // local slab, NIJK over local planes
	bool have_dense = false, have_compressed = false;
	std::vector<oems_coeff_entry> h_table;
	int index_bytes = 0;
	std::vector<float> h_sig[2];
	unsigned sig_len = 0, sig_period = 0;
	ExcHost h_exc[2];
	std::vector<UpmlBoxHost> h_upml;
	std::vector<MurHost> h_mur;
	std::vector<LorHost> h_lor;
	std::vector<RlcHost> h_rlc;
	std::vector<ProbeHost> h_probes;
	unsigned ss_period = 0;
	std::vector<unsigned> ss_pos, ss_dir;
	SsParams pSs{}, pSsF[2];
	bool ss_on = false;

	// device
	float *d_V = nullptr, *d_I = nullptr;
	void* d_idx = nullptr;
	float4* d_tab[10] = {nullptr}; // Evv Evi Pvv Pvvfn Pvvfo Hii Hiv Pii Piifn Piifo
	float *d_flux_v = nullptr, *d_flux_i = nullptr;
	long long flux_floats = 0;
	unsigned* d_numTS = nullptr;
	unsigned numTS_host = 0;
	float* d_sig[2] = {nullptr, nullptr};
	std::vector<void*> allocs; // everything else, freed in the destructor
	uint64_t hbm_bytes = 0;
	uint64_t pml_cells = 0;
	unsigned n_unique = 0;

	StencilParams pE{}, pH{};
	PmlEdgeParams pEdge{};
	bool edge_possible = false; // some UPML box touches the last line of a direction
	int build_edge_list();
	bool has_pml = false;
	MurParams pMur{};
	ExcParams pExc[2]{};
	std::vector<LorDev> lor_dev;
	bool lor_fused = false;          // Lorentz/Drude ADE applied inside the one-pass kernel (kernels_fused_tma.cuh, LOR)
	bool lorentz_fusable() const;
	bool mur_exc_disjoint() const;
	int lor_xmin = 1 << 30, lor_xmax = -1; // x range of the dispersive cells on this engine
	std::string sched_error;         // a schedule that cannot run (reported by iterate)
	std::vector<RlcParams> rlc_dev;
	// probes
	ProbeParams pProbe{};
	unsigned n_values = 0;
	double* d_probe_now = nullptr;
	double* d_series = nullptr;
	unsigned* d_series_ts = nullptr;
	unsigned rec_interval = 0, rec_cap = 0, rec_count = 0;
	std::vector<unsigned> rec_ts;
	bool probes_built = false;
	double* d_energy = nullptr;
	std::vector<DumpHost> dumps;

	// schedule
	int tune_rows = 4, tune_zchunk = 0 /* 0 = auto */, tune_graph = -1;
	int auto_zchunk() const;
	std::vector<std::function<void(cudaStream_t)>> step;
	std::vector<std::string> labels;
	unsigned kernels_per_step = 0;
	uint64_t kernels_launched = 0;
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t graph_exec = nullptr;
	bool use_graph = false;

	// fused one-pass timestep (kernels_fused.cuh): ping-pong field sets (UPML flux is updated in place), one schedule per parity
	float *sV[2] = {nullptr, nullptr}, *sI[2] = {nullptr, nullptr};
	bool pml_disjoint = true;
	ShellParams pShE[2], pShH[2]; // UPML shell launches per parity
	int fused_req = -1; // -1 automatic, 0 two-pass, 1 one-pass
	bool fused_possible = false, fused_active = false;
	int cur() const { return fused_active ? (int)(numTS_host & 1u) : 0; }
	std::vector<std::function<void(cudaStream_t)>> stepf[2];
	std::vector<std::string> labelsf;
	cudaGraph_t graphf[2] = {nullptr, nullptr};
	cudaGraphExec_t graphf_exec[2] = {nullptr, nullptr};
	FusedParams pF[2];
	FusedTmaParams pFT[2]; // the same with the TMA descriptors of the source set (kernels_fused_tma.cuh)
	int tma_req = 1;       // option "tma": stage the inputs of the one-pass kernel through TMA
	bool tma_active = false;
	size_t h2d_bytes = 0;        // every byte copied host -> device by this engine (uploads), counted where the copy is issued
	size_t h2d_index_bytes = 0;  // bytes of operator index copied host -> device
	// UPML boxes updated inside the one-pass kernel ("x slabs", kernels_fused_tma.cuh): index into pE.box, -1 none
	int xs_box[2] = {-1, -1};
	int xslab_req = 2;           // option "xslab": thin UPML boxes at the x ends get their own one-pass kernel: 2 = k_xslab_tma (default), 1 = k_xslab_EH, 0 = shell launches
	XSlabParams pXs[2];          // per parity
	XTmaParams pXt[2];           // the same + TMA descriptors (k_xslab_tma)
	bool xslab_tma = false;
	bool skip_req = true;        // option "skip_shell"
	bool overlap_halo = true;    // option "overlap_halo"
	cudaStream_t side_stream = nullptr;   // parallel branch of the one-pass step graph on z-slab engines
	cudaEvent_t ev_fork[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}, ev_join[2] = {nullptr, nullptr};
	int small_req = -1;          // option "small": one-cell-per-thread two-pass kernels (k_small_E / k_small_H)
	long long small_max_cells = 300000000;
	long long fused_min_cells = 20000000;   // automatic schedule choice (option "fused_min_cells")
	bool small_active = false;
	int skip_active = 0;         // boxes the one-pass kernel skips
	int xs_win[2] = {0, 0};      // first line of the 16-line windows of k_xslab_tma
	int xt_zchunk = 16;
	int make_xslab_maps(int par);
	int nxs = 0;                 // x slabs in pXs
	float* d_flux_v2 = nullptr;  // second voltage-flux set (only the x-slab boxes use it)
	std::vector<int> h_fix_cells;
	void flux_sets_sync(bool to_set0);
	int rebuild_schedule();
	int shell_zchunk = 16;
	int make_tma_maps(int par);
	FixParams pFix[2];
	StencilParams pHtop[2];
	MurParams pMurS[2], pMurD[2];
	ExcParams pExcD[2][2];
	int* d_fix_cells = nullptr;
	long long fix_count = 0;
	int build_fix_list();
	void build_schedule_fused();
	int set_fused_active(int req);
	bool fused_auto_choice() const;
	float *peer_lo_Vs[2] = {nullptr, nullptr}, *peer_hi_Is[2] = {nullptr, nullptr};

	// multi-GPU
	Engine *peer_lo = nullptr, *peer_hi = nullptr;
	float *peer_lo_V = nullptr, *peer_hi_I = nullptr; // mapped neighbour field bases
	unsigned *peer_lo_flagE = nullptr, *peer_hi_flagH = nullptr;
	unsigned *d_flagE = nullptr, *d_flagH = nullptr, *d_halo_cnt = nullptr, *d_halo_err = nullptr;
	long long peer_lo_comp = 0, peer_hi_comp = 0, peer_lo_ghostE_off = 0, peer_hi_ghostH_off = 0;
	bool peers_linked = false;
	int halo_timeout_s = 600;
	long long halo_timeout_cycles() const;
	// complete ghost planes for the readout (k_ghost_push): the neighbours' other field bases and the flags
	float *peer_lo_Is[2] = {nullptr, nullptr}, *peer_hi_Vs[2] = {nullptr, nullptr};
	unsigned *peer_lo_flags = nullptr, *peer_hi_flags = nullptr; // the neighbours' flag blocks (index: FLAG_*)
	enum { FLAG_E = 0, FLAG_H = 1, FLAG_CNT = 2, FLAG_ERR = 4, FLAG_G_LO = 5, FLAG_G_HI = 6, FLAG_ACK_LO = 7, FLAG_ACK_HI = 8, FLAG_GCNT = 9, FLAG_WORDS = 16 };
	unsigned ghost_seq = 0;       // exchanges issued so far
	unsigned ghost_ts = 0;        // numTS of the last one
	bool ghost_open = false;      // the neighbours have not been told yet that this slab is done reading
	int ghosts_for_readout();
public:
	int exchange_ghosts();
	int release_ghosts();
private:
	std::vector<void*> ipc_opened;

	template <typename T> T* dalloc(size_t n, bool zero = true);
	template <typename T> T* upload(const std::vector<T>& v);
	int compress_dense(std::vector<uint32_t>& index32);
	int build_tables_and_index(const std::vector<uint32_t>& index32);
	int build_pml();
	int build_mur();
	int build_exc();
	int build_lorentz();
	int build_rlc();
	int build_probes();
	void build_schedule();
	void launch_probes(double* dst);
	void mark_edge_dirty();
public:
	int set_option(const char* key, long long value);
	int get_option(const char* key, long long* value);
private:
	bool edge_dirty = false;
	bool owned(unsigned z) const { return z >= zb && z < ze; }
	bool held(unsigned z) const { return (int)z >= z0 && (int)z < z0 + nzl; }
	long long cell_off(unsigned x, unsigned y, unsigned z) const
	{
		return (long long)((int)z - z0) * plane + (long long)y * pitch + x;
	}
	int check(cudaError_t e, const char* what);
};
