/*
 * openems_b200.h -- C ABI of the B200-native openEMS FDTD engine (libopenems_b200.so).
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json: everything the
 * reference's Operator / Engine / Engine_Extension / Engine_Interface_Base classes need from
 * a device engine, as plain C.  The reference has no C ABI or plugin loader for engines
 * (engines are C++ classes created by Operator::CreateEngine, FDTD/operator.h:56), so each
 * entry point below names the reference interface it replaces; INTEGRATION.md shows the
 * Engine_CUDA / Operator_CUDA C++ classes that bind these calls inside openEMS.
 *
 * Conventions
 *  - every call returns 0 on success, non-zero on error; oems_cuda_last_error() gives the text.
 *    The reference reports errors by cerr + return code / exit (openems.cpp:1133-1344); no
 *    exception crosses this boundary.
 *  - all pointers are HOST pointers, copied during the call; the caller keeps ownership.
 *  - dense arrays use the reference's ArrayNIJK order [n][i][j][k], k (z) fastest
 *    (tools/arraylib/array_nijk.h) unless stated otherwise.
 *  - one handle per host thread (the reference calls IterateTS and all Processing from one
 *    thread, openems.cpp:1427-1476); calls on one handle must not overlap.
 *  - there is no CPU fallback: create() fails if no CUDA device is usable.
 */
#ifndef OPENEMS_B200_H
#define OPENEMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
/* The library is built with -fvisibility=hidden: only the C entry points declared here are exported.  Its internal
   C++ classes (one of them is called Engine, like FDTD/engine.h's) must never be visible to, or be interposed by,
   the host application's symbols. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct oems_cuda_engine oems_cuda_engine;

#define OEMS_ABI_VERSION 1

/* one de-duplicated coefficient tuple of the device operator.  Replaces the
   SSE_coeff key of Operator_SSE_Compressed (FDTD/operator_sse_compressed.h:77-85), re-keyed
   per cell and extended by the UPML auxiliary coefficients so that one index serves the
   stencil and the PML (FDTD/extensions/operator_ext_upml.h:108-113). 32 floats = 128 B. */
typedef struct oems_coeff_entry {
	float vv[3], vi[3], ii[3], iv[3];
	float pml;                 /* != 0: cell lies inside a UPML box, aux coefficients valid */
	float pml_vv[3], pml_vvfn[3], pml_vvfo[3];
	float pml_ii[3], pml_iifn[3], pml_iifo[3];
	float reserved;
} oems_coeff_entry;

int         oems_cuda_abi_version(void);
/* text of the last error on this handle (or of the last failed create() when h == NULL) */
const char* oems_cuda_last_error(const oems_cuda_engine* h);

/* ---- life cycle: Engine::New / ~Engine (FDTD/engine.cpp:27-49); numLines as
   Operator::GetNumberOfLines(n,true).  device < 0 selects the current CUDA device. */
int oems_cuda_create(unsigned nx, unsigned ny, unsigned nz, int device, oems_cuda_engine** out);
int oems_cuda_destroy(oems_cuda_engine* h);

/* ---- z-slab sharding (multi-GPU; template: FDTD/engine_mpi.cpp:84-210, openems_fdtd_mpi.cpp:201-299).
   Must be called before any upload.  The handle then holds global planes
   [z_begin - (z_begin>0), z_end + (z_end<nz)) : owned planes plus one ghost plane per
   interface; all positions passed to other calls stay GLOBAL.  Single GPU: never call it. */
int oems_cuda_set_slab(oems_cuda_engine* h, unsigned z_begin, unsigned z_end);

/* ---- operator upload (once) */
/* dense coefficients as held by Operator (GetVV/GetVI/GetII/GetIV, FDTD/operator.h:215-218):
   the library re-keys them per cell (replaces Operator_SSE_Compressed::CompressOperator,
   FDTD/operator_sse_compressed.cpp:114-175). */
int oems_cuda_set_operator_dense(oems_cuda_engine* h, const float* vv, const float* vi,
                                 const float* ii, const float* iv);
/* already compressed operator: table[n_unique] + per-cell index [nz][ny][nx] (x fastest),
   index_bytes 2 or 4.  Used when the dense arrays do not fit the host (1024^3 and up).
   A pageable index is staged and copied before the call returns; a PAGE-LOCKED index
   (cudaHostRegister / cudaMallocHost, e.g. oems_synth_pin) is sent by one asynchronous DMA that
   overlaps the allocations of oems_cuda_finalize: keep that buffer alive until finalize returns. */
int oems_cuda_set_operator_compressed(oems_cuda_engine* h, unsigned n_unique,
                                      const oems_coeff_entry* table, const void* index,
                                      int index_bytes);
/* the same, for operators that repeat along z (uniform or piecewise-uniform meshes, layered structures):
   the per-cell index as n_planes unique xy planes [n_planes][ny][nx] plus one plane id per mesh line in z;
   the engine expands it on the device.  C5 (1024^3, PML_8): 20 planes = 40 MB instead of 2.1 GB over PCIe. */
int oems_cuda_set_operator_planes(oems_cuda_engine* h, unsigned n_unique, const oems_coeff_entry* table,
                                  unsigned n_planes, const void* planes, const unsigned* plane_of_z,
                                  int index_bytes);
/* Excitation::GetVoltageSignal/GetCurrentSignal/GetLength/GetSignalPeriod (FDTD/excitation.h);
   period_ts = int(GetSignalPeriod()/GetTimestep()) or 0 (engine_ext_excitation.cpp:43-45) */
int oems_cuda_set_signal(oems_cuda_engine* h, const float* sig_volt, const float* sig_curr,
                         unsigned length, unsigned period_ts);
/* Operator_Ext_Excitation lists Volt_index/Volt_dir/Volt_amp/Volt_delay and Curr_*
   (FDTD/extensions/operator_ext_excitation.h:81-94); idx3 is [3][count]. */
int oems_cuda_add_excitation(oems_cuda_engine* h, int is_curr, unsigned count,
                             const unsigned* idx3, const unsigned* dir, const float* amp,
                             const unsigned* delay);
/* one Operator_Ext_UPML box: m_StartPos, m_numLines and the six coefficient arrays
   vv, vvfn, vvfo, ii, iifn, iifo, each ArrayNIJK local to the box
   (FDTD/extensions/operator_ext_upml.h:96-113).  With a compressed operator pass NULL
   coefficient pointers: the aux coefficients then come from the table entries. */
int oems_cuda_add_upml(oems_cuda_engine* h, const unsigned start[3], const unsigned nlines[3],
                       const float* vv, const float* vvfn, const float* vvfo,
                       const float* ii, const float* iifn, const float* iifo);
/* one Operator_Ext_Mur_ABC plane: m_ny, m_LineNr, m_LineNr_Shift, m_numLines, the two
   coefficient arrays ArrayIJ [nyP][nyPP] (FDTD/extensions/operator_ext_mur_abc.h:93-103) and
   the engine's m_start_TS (engine_ext_mur_abc.cpp:44-60).  Call in the reference's insertion
   order xmin..zmax (openems.cpp:388-396). */
int oems_cuda_add_mur(oems_cuda_engine* h, int ny, unsigned line_nr, unsigned line_nr_shift,
                      const unsigned nlines[2], const float* coeff_nyP, const float* coeff_nyPP,
                      unsigned start_ts);
/* one dispersion order of Operator_Ext_LorentzMaterial / ConductingSheet: m_LM_pos[o],
   v_int/v_ext/v_Lor/i_int/i_ext/i_Lor[o] as [3][count]; NULL = that ADE is off
   (FDTD/extensions/operator_ext_lorentzmaterial.h:53-62, operator_ext_dispersive.h).
   Call once per order, in order. */
int oems_cuda_add_lorentz(oems_cuda_engine* h, unsigned count, const unsigned* pos3,
                          const float* v_int, const float* v_ext, const float* v_lor,
                          const float* i_int, const float* i_ext, const float* i_lor);
/* Operator_Ext_LumpedRLC: v_RLC_dir, v_RLC_pos[3][count] and the nine coefficient arrays
   (FDTD/extensions/operator_ext_lumpedRLC.h:65-82) */
int oems_cuda_add_rlc(oems_cuda_engine* h, unsigned count, const int* dir, const unsigned* pos3,
                      const float* ilv, const float* i2v, const float* vvd, const float* vv2,
                      const float* vj1, const float* vj2, const float* ib0, const float* b1,
                      const float* b2);
/* Operator_Ext_SteadyState (created by the driver for periodic excitations, openems.cpp:1206-1234):
   m_TS_period and the E probe list m_E_probe_pos / m_E_probe_dir
   (FDTD/extensions/operator_ext_steadystate.h).  The probe voltages are recorded on the device
   every timestep; oems_cuda_steadystate_check evaluates Engine_Ext_SteadyState::Apply2Voltages
   (engine_ext_steadystate.cpp:50-107) for the last completed period and returns what
   GetLastDiff() would (1 until two periods have passed). */
int oems_cuda_add_steadystate(oems_cuda_engine* h, unsigned period_ts, unsigned count,
                              const unsigned* pos3, const unsigned* dir);
int oems_cuda_steadystate_check(oems_cuda_engine* h, double* last_diff, unsigned* n_checks);
/* Operator_Ext_TFSF (plane-wave excitation, FDTD/extensions/operator_ext_tfsf.h): m_Start, m_Stop,
   m_ActiveDir[n][l] as active6[2n+l], and the per-face tables m_VoltDelay / m_VoltDelayDelta /
   m_VoltAmp / m_CurrDelay / m_CurrDelayDelta / m_CurrAmp [n][l][c] passed as 12 pointers each, index
   (n*2+l)*2+c, NULL for inactive faces; each table has numLines[nP]*numLines[nPP] entries.  The
   engine runs Engine_Ext_TFSF::DoPostVoltageUpdates / DoPostCurrentUpdates
   (engine_ext_tfsf.cpp:36-215) on the device, in the reference's update order and C++ arithmetic
   types; the delay lookup is evaluated per update (same expression as m_DelayLookup). */
int oems_cuda_set_tfsf(oems_cuda_engine* h, const unsigned* start3, const unsigned* stop3, const int* active6,
                       const unsigned* const* volt_delay, const float* const* volt_delay_delta, const float* const* volt_amp,
                       const unsigned* const* curr_delay, const float* const* curr_delay_delta, const float* const* curr_amp);
/* Operator_Ext_Absorbing_BC (local absorbing sheet, one per CSXCAD primitive, openems.cpp:411-441;
   FDTD/extensions/operator_ext_absorbing_bc.h:94-113): m_ny, m_sheetX0 / m_sheetX1 (mesh indices),
   m_normalSignPositive, m_ABCtype (1 MUR_1ST, 2 MUR_1ST_SA) and the ArrayIJ coefficient tables
   m_K1_nyP / m_K1_nyPP [numLines0][numLines1] (+ m_K2_* for type 2).  The six hooks of
   Engine_Ext_Absorbing_BC (engine_ext_absorbing_bc.cpp:108-366) run on the device in the reference's
   order (sheets are the last extensions inserted: first in the post/apply lists). */
int oems_cuda_add_absorbing_sheet(oems_cuda_engine* h, int ny, const unsigned* x0, const unsigned* x1, int normal_positive, int type,
                                  const float* K1_nyP, const float* K1_nyPP, const float* K2_nyP, const float* K2_nyPP);
/* ends the upload: compresses, moves everything to HBM, fixes the extension schedule in the
   order of Engine::SortExtensionByPriority (FDTD/engine.cpp:87-98) and captures the
   per-timestep CUDA graph.  Replaces Engine::Init (engine.cpp:51-59). */
int oems_cuda_finalize(oems_cuda_engine* h);

/* ---- time stepping: Engine::IterateTS / GetNumberOfTimesteps (FDTD/engine.h:48-50).
   iterate() only enqueues; reads synchronise. */
int oems_cuda_iterate(oems_cuda_engine* h, unsigned n_ts);
/* iterate + device time of the burst in ms (CUDA events on the engine's stream), synchronises */
int oems_cuda_iterate_timed(oems_cuda_engine* h, unsigned n_ts, double* elapsed_ms);
int oems_cuda_sync(oems_cuda_engine* h);
int oems_cuda_num_ts(oems_cuda_engine* h, unsigned* ts);
int oems_cuda_reset(oems_cuda_engine* h); /* fields, extension state and numTS back to 0 */

/* ---- readout (replaces the per-cell virtual Engine::GetVolt/GetCurr walks of L3/L4) */
/* Engine_Interface_FDTD::CalcVoltageIntegral (FDTD/engine_interface_fdtd.cpp:206-232): fp64
   sum of fp32 edge voltages in the reference's order.  Values: 1. */
int oems_cuda_add_probe_voltage(oems_cuda_engine* h, const unsigned start[3],
                                const unsigned stop[3], int* probe_id);
/* ProcessCurrent::CalcIntegral (Common/processcurrent.cpp:96-171): fp32 loop sum, same order,
   with the m_start_inside / m_stop_inside flags of Processing.  Values: 1. */
int oems_cuda_add_probe_current(oems_cuda_engine* h, const unsigned start[3],
                                const unsigned stop[3], int norm_dir, const int start_inside[3],
                                const int stop_inside[3], int* probe_id);
/* ProcessFieldProbe (Common/processfieldprobe.cpp:77-92): raw V (is_H=0) or I (is_H=1) of the
   three components at one node; the caller divides by the edge length like
   GetRawField/GetRawDualField (engine_interface_fdtd.cpp:263-268,136-141).  Values: 3. */
int oems_cuda_add_probe_field(oems_cuda_engine* h, int is_H, const unsigned pos[3], int* probe_id);
int oems_cuda_num_probe_values(oems_cuda_engine* h, unsigned* n_values);
/* all probe values now, in probe_id order (Processing::Process between bursts) */
int oems_cuda_read_probes(oems_cuda_engine* h, double* out);
/* device-side recording: every `interval` timesteps (numTS % interval == 0, the cadence of
   Processing::CheckTimestep, Common/processing.cpp:82-105) all probe values are appended to
   a device series, read back later in one copy.  interval 0 switches recording off. */
int oems_cuda_record_probes(oems_cuda_engine* h, unsigned interval, unsigned max_samples);
/* out: [n_samples][n_values]; ts_out (optional): numTS of each sample */
int oems_cuda_read_probe_series(oems_cuda_engine* h, double* out, unsigned* ts_out,
                                unsigned capacity_samples, unsigned* n_samples);
/* Engine_Interface_FDTD::CalcFastEnergy (engine_interface_fdtd.cpp:302-347) */
int oems_cuda_energy(oems_cuda_engine* h, double* energy);

/* Measurement aids (bench.py).  oems_cuda_fill_fields: deterministic pre-fill of E and H of the held
   planes, a function of the global cell index and the seed only ((hash mod 2^16 - 2^15) * 1e-6; H scaled by 1/Z0,
   zero on the last line of each direction), so that a timed window does not run on an all-zero domain.
   oems_cuda_field_digest: order-independent 64-bit digest (sum mod 2^64 over the OWNED cells of a hash of global
   index and bit pattern): the digests of z-slab engines add up to the single-GPU digest of the same state. */
int oems_cuda_fill_fields(oems_cuda_engine* h, unsigned long long seed);
int oems_cuda_field_digest(oems_cuda_engine* h, int is_curr, unsigned long long* out);

/* ProcessFields::CalcField box dump (Common/processfields.cpp:283-409) with the
   NO/NODE/CELL interpolation of engine_interface_fdtd.cpp:63-124,150-204, gathered and
   interpolated on the device.  px/py/pz are the posLines index lists; edge_len[n] /
   dual_edge_len[n] are Operator::GetEdgeLength(n,pos,false/true) per line of direction n
   (lengths nx, ny, nz).  Output order {3,nz,ny,nx}, x fastest, float
   (tools/hdf5_file_writer.cpp:286-302). interp: 0 none, 1 node, 2 cell. */
int oems_cuda_add_dump(oems_cuda_engine* h, int is_H, int interp, unsigned nx, unsigned ny,
                       unsigned nz, const unsigned* px, const unsigned* py, const unsigned* pz,
                       const double* edge_len[3], const double* dual_edge_len[3], int* dump_id);
/* computes the dump on the device at the current numTS and copies it to `out`
   (asynchronously into pinned staging, then to the caller's buffer) */
int oems_cuda_read_dump(oems_cuda_engine* h, int dump_id, float* out);
/* the same without stalling the time loop (ProcessFieldsTD::Process, Common/processfields_td.cpp:50-91, is the
   consumer): the dump is evaluated on the device at the current numTS, the copy into `pinned_out` (page-locked,
   e.g. from oems_cuda_host_alloc; 3*nx*ny*nz floats) runs on a second stream while oems_cuda_iterate goes on.
   oems_cuda_wait(ticket) returns when that copy has landed.  One copy per dump box may be outstanding; issuing
   the next one for the same box orders itself after the previous copy on the device (no host wait). */
int oems_cuda_read_dump_async(oems_cuda_engine* h, int dump_id, float* pinned_out, long long* ticket);
int oems_cuda_wait(oems_cuda_engine* h, long long ticket);
/* page-locked host memory for asynchronous read-outs */
int oems_cuda_host_alloc(size_t bytes, void** out);
int oems_cuda_host_free(void* p);

/* ProcessFieldsFD (Common/processfields_fd.cpp:40-107): running DFT of a field dump, kept on the
   device.  oems_cuda_add_fd_dump attaches n_freq complex<float> accumulators (zero) to a dump
   made by oems_cuda_add_dump.  oems_cuda_fd_accumulate is ProcessFieldsFD::Process for one
   sample: it evaluates the dump at the current timestep on the device (no D2H) and does
   field_fd[f] += field_td * w[f] with w[f] = {re, im} of the reference's exp_jwt_2_dt
   (processfields_fd.cpp:84-86), which the caller computes exactly as the reference does.
   oems_cuda_read_fd (PostProcess / DumpFDData) copies out[n_freq][3][nz][ny][nx][2] -- the block
   HDF5_File_Writer::WriteVectorField(name, complex<float>****, ...) splits into _real / _imag. */
int oems_cuda_add_fd_dump(oems_cuda_engine* h, int dump_id, unsigned n_freq, int* fd_id);
int oems_cuda_fd_accumulate(oems_cuda_engine* h, int fd_id, const float* weights_re_im);
int oems_cuda_read_fd(oems_cuda_engine* h, int fd_id, float* out_re_im, unsigned* n_samples);

/* ProcessModeMatch (Common/processmodematch.cpp:71-266): waveguide-port mode matching.  start3 /
   stop3 = the surface AFTER InitProcess has sorted it and excluded the boundaries (lines 86-104),
   ny its normal; dist0 / dist1 = the normalised m_ModeDist[0/1][posP][posPP] (fparser evaluation and
   normalisation stay on the host, lines 119-190), area = Op->GetNodeArea(ny,pos,dualMesh) of the same
   points; is_H = m_ModeFieldType.  oems_cuda_read_mode_match is CalcMultipleIntegrals (lines
   222-266) at the current timestep: out2 = {value, value^2/purity}, node-interpolated fields, summed
   in the reference's loop order. */
int oems_cuda_add_mode_match(oems_cuda_engine* h, int is_H, int ny, const unsigned* start3, const unsigned* stop3,
                             const double* dist0, const double* dist1, const double* area,
                             const double* const edge_len[3], const double* const dual_edge_len[3], int* id);
int oems_cuda_read_mode_match(oems_cuda_engine* h, int id, double* out2);

/* slow path for unknown callers: Engine::GetVolt/SetVolt/GetCurr/SetCurr (FDTD/engine.h:55-101) */
int oems_cuda_get_field(oems_cuda_engine* h, int is_curr, unsigned n, unsigned x, unsigned y,
                        unsigned z, float* value);
int oems_cuda_set_field(oems_cuda_engine* h, int is_curr, unsigned n, unsigned x, unsigned y,
                        unsigned z, float value);
/* whole field, ArrayNIJK order over the planes this handle holds (all of them on one GPU) */
int oems_cuda_get_fields(oems_cuda_engine* h, int is_curr, float* out_nijk);
int oems_cuda_set_fields(oems_cuda_engine* h, int is_curr, const float* in_nijk);
/* UPML flux state of box b (Engine_Ext_UPML::volt_flux / curr_flux), box-local ArrayNIJK */
int oems_cuda_get_upml_flux(oems_cuda_engine* h, int box, int is_curr, float* out_nijk);

/* ---- statistics for DESIGN/bench: de-duplicated entries, index width, bytes in HBM,
   kernels launched so far, kernels per timestep */
typedef struct oems_cuda_stats {
	unsigned n_unique;
	int      index_bytes;
	uint64_t hbm_bytes;
	uint64_t kernels_launched;
	unsigned kernels_per_step;
	unsigned pml_cells_lo, pml_cells_hi; /* 64-bit count split */
	int      uses_graph;
} oems_cuda_stats;
int oems_cuda_get_stats(oems_cuda_engine* h, oems_cuda_stats* out);
/* tuning knobs (0 keeps the default): block rows and z-chunk of the stencil kernels, graph on/off */
int oems_cuda_set_tuning(oems_cuda_engine* h, int block_rows, int z_chunk, int use_graph);

/* named integer options (all variants give identical results; they select kernels):
   "fused" = 0 / 1 / -1: two-pass (in place), one-pass (E and H in one kernel, ping-pong field sets) or automatic
     choice of the timestep schedule.  One-pass needs a hook set without lumped RLC, Lorentz/Drude cells only where no
     other hook reads them between the stencil and Apply2Voltages, disjoint UPML boxes and memory for
     the second field set; automatic picks it from "fused_min_cells" (default 20 M) cells per GPU.
   "small" = 0 / 1 / -1: two-pass half-steps with one cell per thread (k_small_E / k_small_H) instead of the float4 /
     z-march kernels; automatic below "small_max_cells" (default 300 M) cells per GPU.
   "tma" = 1 / 0: the one-pass kernel stages its inputs in shared memory through TMA bulk tensor copies (default) or
     loads them into registers itself (k_fused_EH).
   "xslab" = 2 / 1 / 0: UPML boxes that are thin in x and sit at the x ends of the mesh are updated by the
     TMA-staged one-pass window kernel k_xslab_tma (default), by k_xslab_EH, or by the shell launches.
   "skip_shell" = 1 / 0: the one-pass kernels skip the planes / rows at the mesh ends that consist of UPML cells only
     (default) or pass them through.
   "pdl" = 0 / 1: programmatic dependent launch between the kernels of a timestep (default off: no gain measured).
   "halo_timeout_s": seconds a z-slab waits for its neighbour's halo before the device traps (default 600).
   "xslab_zchunk", "shell_zchunk": planes a block of those kernels marches (tuning aids). */
int oems_cuda_set_option(oems_cuda_engine* h, const char* key, long long value);
/* reports what is ACTIVE: "fused", "tma", "small", "pdl": 0 / 1; "xslab": boxes on the window kernel; "skip_shell": ends
   of the mesh that are skipped; "onepass_rows" / "onepass_planes": rows / local planes the one-pass kernel works on;
   "h2d_bytes": bytes copied host -> device by this engine so far. */
int oems_cuda_get_option(oems_cuda_engine* h, const char* key, long long* value);

/* measurement aid: runs n_ts timesteps without the graph and returns the average duration in
   ms of every kernel of the per-timestep schedule, timed with CUDA events on the engine's own
   stream; oems_cuda_schedule_label(i) names entry i ("update_E", "update_H", "mur_pre", ...) */
int oems_cuda_time_schedule(oems_cuda_engine* h, unsigned n_ts, double* ms_out, unsigned capacity,
                            unsigned* n_entries);
const char* oems_cuda_schedule_label(oems_cuda_engine* h, unsigned i);

/* ---- multi-GPU halo exchange over NVLink peer memory (one process per GPU).
   Each rank exports IPC handles of its field buffers, gathers its neighbours' with
   torch.distributed / any host channel, and opens them here; the halo planes are then written
   straight into the neighbour's ghost plane by the boundary kernels (no host staging, no
   collective), with device-side flags ordering the steps.  See DESIGN.md "multi-GPU". */
#define OEMS_IPC_BYTES 512
int oems_cuda_export_ipc(oems_cuda_engine* h, unsigned char out[OEMS_IPC_BYTES]);
int oems_cuda_open_peers(oems_cuda_engine* h, const unsigned char* lower /*OEMS_IPC_BYTES or NULL*/,
                         const unsigned char* upper);
/* same-process variant (tests, or one process driving several GPUs) */
int oems_cuda_link_peers(oems_cuda_engine* h, oems_cuda_engine* lower, oems_cuda_engine* upper);

/* ---- readout on z-slab engines (reference template: every MPI rank of Engine_MPI processes the part of a probe /
   dump box that lies in its sub-domain, FDTD/openems_fdtd_mpi.cpp:201-299, and holds complete neighbour planes,
   FDTD/engine_mpi.cpp:84-182).
   The time loop only exchanges the tangential E / H components the stencil needs.  The interpolating readers
   (field dumps, FD dumps, mode matching: engine_interface_fdtd.cpp:63-124,150-204) also read the normal components
   and E of the plane below / H of the plane above, so before such a readout every slab completes its neighbours'
   ghost planes: oems_cuda_exchange_ghosts pushes all components of E and H of its lowest / highest owned plane
   through the peer mappings and waits (on the device) for the neighbours' planes.  It only enqueues, is a no-op
   on an engine without neighbours, and is called implicitly by oems_cuda_read_dump(_async), oems_cuda_fd_accumulate
   and oems_cuda_read_mode_match(_raw); one exchange serves all readouts of the same timestep.  COLLECTIVE: every
   slab has to reach the same readout at the same timestep.  A single host thread that drives several slabs must call
   oems_cuda_exchange_ghosts on all of them before the first call that blocks (oems_cuda_read_dump, _wait, ...).
   The next oems_cuda_iterate (or an explicit oems_cuda_release_ghosts) tells the neighbours that this slab is
   done reading and waits for theirs before the time loop writes into their ghost planes again. */
int oems_cuda_exchange_ghosts(oems_cuda_engine* h);
int oems_cuda_release_ghosts(oems_cuda_engine* h);
/* a z-slab engine evaluates the z lines of a dump box it owns: entries [first, first + n) of the pz list given to
   oems_cuda_add_dump (which must ascend); oems_cuda_read_dump / _read_fd return 3*nx*ny*n values, and the host
   concatenates the slabs along z.  On a single-GPU engine first = 0, n = nz. */
int oems_cuda_dump_own_range(oems_cuda_engine* h, int dump_id, unsigned* first, unsigned* n);
/* mode matching over the planes this engine owns: out3 = {value, value^2/purity, purity}.  Slab partial results are
   put together on the host: value = sum of value, purity = sum of purity, result {value, value^2/purity}. */
int oems_cuda_read_mode_match_raw(oems_cuda_engine* h, int id, double* out3);
/* steady-state detection on z-slab engines: every slab records the probes it owns (the order of the list given to
   oems_cuda_add_steadystate is kept) and the energy of its planes.  oems_cuda_steadystate_raw returns
   info2 = {checks done, timestep of the last}, energy4 = {scratch, scratch, previous period, current period},
   snap[2*period][count] and the local probe count; the host adds the energies, concatenates the records (count
   fastest) and evaluates Engine_Ext_SteadyState::Apply2Voltages (engine_ext_steadystate.cpp:62-106) with
   oems_cuda_steadystate_eval (host arithmetic, no engine needed). */
int oems_cuda_steadystate_raw(oems_cuda_engine* h, unsigned* info2, double* energy4, double* snap, unsigned capacity,
                              unsigned* count);
int oems_cuda_steadystate_eval(unsigned period_ts, unsigned count, const unsigned* info2, const double* energy4,
                               const double* snap, double* last_diff);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
