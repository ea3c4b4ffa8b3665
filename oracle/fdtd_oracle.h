/*
 * fdtd_oracle.h -- CPU restatement of the openEMS FDTD hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the B200 engine in openems_b200/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path never links or calls anything in oracle/.
 *
 * Every function restates a reference function; the citation (file:line relative to the
 * openEMS tree) is given beside each prototype and again at the definition.
 *
 * Parity pinning: the reference holds no golden vectors for this path
 * (SURVEY.md section 4 / 8c).  The oracle is pinned by
 *   (1) the analytic PEC-cavity resonances of TESTSUITE/combinedtests/cavity.m:24-32,
 *   (2) the cross-engine bit-equality rule of TESTSUITE/enginetests/cavity.m:155
 *       (scalar restatement == sse-compressed restatement, bit for bit),
 *   (3) the probe == dump rule of TESTSUITE/probes/fieldprobes.m:34,
 *   (4) closed forms on a uniform mesh (dT, vi, iv).
 * Geometry-driven coefficients (CSXCAD / fparser, not vendored) are restated from their
 * call sites only: "parity unpinned" for that part, see DESIGN.md.
 *
 * Array layout everywhere in this file: ArrayNIJK  [n][i][j][k], k (z) fastest
 * (tools/arraylib/array_nijk.h), exactly what Engine (FDTD/engine.cpp) uses.
 */
#ifndef FDTD_ORACLE_H
#define FDTD_ORACLE_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_sim orc_sim;

/* boundary condition codes, openems.cpp:837-975 / operator.h (m_BC): */
enum { ORC_BC_NONE = -1, ORC_BC_PEC = 0, ORC_BC_PMC = 1, ORC_BC_MUR = 2, ORC_BC_PML = 3 };
/* CSPropExcitation types used by operator_ext_excitation.cpp:176-231 */
enum { ORC_EXC_E_SOFT = 0, ORC_EXC_E_HARD = 1, ORC_EXC_H_SOFT = 2, ORC_EXC_H_HARD = 3 };

/* ---- mesh / geometry (stand-in for CSXCAD + Operator::SetupCSXGrid, operator.cpp:793-820) */
orc_sim* orc_create(const unsigned nlines[3], const double* x, const double* y, const double* z,
                    double grid_delta);
void     orc_destroy(orc_sim* s);
void     orc_set_bc(orc_sim* s, const int bc[6], const unsigned pml_size[6]);
void     orc_set_background(orc_sim* s, double epsR, double mueR, double kappa, double sigma);
void     orc_set_mur_phase_velocity(orc_sim* s, double v_phase);
void     orc_set_timestep(orc_sim* s, double forced_dT, double factor);
/* geometry primitives are axis-aligned boxes in drawing units, containment inclusive;
   the highest priority wins, ties go to the box added later. Returns the property id. */
int orc_add_material(orc_sim* s, int prio, const double start[3], const double stop[3],
                     double epsR, double mueR, double kappa, double sigma);
int orc_add_metal(orc_sim* s, int prio, const double start[3], const double stop[3]);
/* Drude/Lorentz material: per pole o<order the plasma frequency, relaxation time and
   Lorentz pole frequency for eps and mue (operator_ext_lorentzmaterial.cpp:240-262). */
int orc_add_lorentz(orc_sim* s, int prio, const double start[3], const double stop[3],
                    double epsR, double mueR, double kappa, double sigma, int order,
                    const double* eps_fp, const double* eps_tau, const double* eps_flor,
                    const double* mue_fp, const double* mue_tau, const double* mue_flor);
int orc_add_excitation(orc_sim* s, int prio, const double start[3], const double stop[3],
                       int exc_type, const double vec[3], double delay_s);
/* parallel RC lumped element folded into vv/vi, Operator::Calc_LumpedElements operator.cpp:1586-1763 */
int orc_add_lumped_rc(orc_sim* s, const double start[3], const double stop[3], int dir,
                      double R, double C, int caps);
/* series/parallel RLC engine extension given directly by its 9 coefficient arrays
   (engine_ext_lumpedRLC.cpp:83-142; names as operator_ext_lumpedRLC.h:65-82) */
int orc_add_rlc_raw(orc_sim* s, unsigned count, const int* dir, const unsigned* pos /*[3][count]*/,
                    const float* ilv, const float* i2v, const float* vvd, const float* vv2,
                    const float* vj1, const float* vj2, const float* ib0, const float* b1,
                    const float* b2);

/* steady-state detection: Engine_Ext_SteadyState engine_ext_steadystate.cpp:50-107 with the
   probe set of openems.cpp:1206-1234; pos3 is [3][count]. */
int    orc_add_steadystate(orc_sim* s, unsigned period_ts, unsigned count, const unsigned* pos3, const int* dir);
double orc_steadystate_last_diff(const orc_sim* s);

/* ---- excitation signal, FDTD/excitation.cpp:150-276 */
void orc_set_excite_gauss(orc_sim* s, double f0, double fc);
void orc_set_excite_sinus(orc_sim* s, double f0);
void orc_set_excite_dirac(orc_sim* s, double fmax);
void orc_set_excite_step(orc_sim* s, double fmax);

/* ---- operator build: Operator::CalcECOperator operator.cpp:986-1097 + extension builds,
        then Engine::Init engine.cpp:51-98. returns 0 on success. */
int orc_build(orc_sim* s, unsigned max_ts);

/* ---- operator results */
double       orc_dT(const orc_sim* s);
unsigned     orc_nyquist(const orc_sim* s);
const float* orc_coeff(const orc_sim* s, int which /*0 vv 1 vi 2 ii 3 iv*/);
unsigned     orc_signal_length(const orc_sim* s);
const float* orc_signal(const orc_sim* s, int is_curr);
unsigned     orc_signal_period_ts(const orc_sim* s); /* 0 = not periodic */
unsigned     orc_exc_count(const orc_sim* s, int is_curr);
/* out arrays: idx[3][count] (unsigned), dir[count] (unsigned), amp[count], delay[count] */
void orc_exc_get(const orc_sim* s, int is_curr, unsigned* idx, unsigned* dir, float* amp,
                 unsigned* delay);
int  orc_upml_count(const orc_sim* s);
void orc_upml_box(const orc_sim* s, int b, unsigned start[3], unsigned nlines[3]);
const float* orc_upml_coeff(const orc_sim* s, int b, int which /*0 vv 1 vvfn 2 vvfo 3 ii 4 iifn 5 iifo*/);
int  orc_mur_count(const orc_sim* s);
void orc_mur_info(const orc_sim* s, int m, int* ny, int* top, unsigned* line, unsigned* shift,
                  unsigned nlines[2], unsigned* start_ts);
const float* orc_mur_coeff(const orc_sim* s, int m, int which /*0 nyP 1 nyPP*/);
int  orc_lorentz_order(const orc_sim* s);
unsigned orc_lorentz_count(const orc_sim* s, int o);
/* flags bit0 volt_ADE_On bit1 curr_ADE_On bit2 volt_Lor_On bit3 curr_Lor_On */
int  orc_lorentz_flags(const orc_sim* s, int o);
const unsigned* orc_lorentz_pos(const orc_sim* s, int o, int n);
/* which: 0 v_int 1 v_ext 2 v_Lor 3 i_int 4 i_ext 5 i_Lor ; NULL if not allocated */
const float* orc_lorentz_coeff(const orc_sim* s, int o, int which, int n);

/* ---- engine: Engine::IterateTS engine.cpp:267-286 with the hook order of engine.cpp:224-265 */
void     orc_iterate(orc_sim* s, unsigned n_ts);
unsigned orc_num_ts(const orc_sim* s);
float*   orc_volt(orc_sim* s);
float*   orc_curr(orc_sim* s);
const float* orc_upml_flux(const orc_sim* s, int b, int is_curr);
void     orc_reset_fields(orc_sim* s);

/* ---- readout */
/* Engine_Interface_FDTD::CalcVoltageIntegral engine_interface_fdtd.cpp:206-232 */
double orc_voltage_integral(const orc_sim* s, const unsigned start[3], const unsigned stop[3]);
/* ProcessCurrent::CalcIntegral Common/processcurrent.cpp:96-171 */
double orc_current_integral(const orc_sim* s, const unsigned start[3], const unsigned stop[3],
                            int norm_dir, const int start_inside[3], const int stop_inside[3]);
/* Engine_Interface_FDTD::GetRawField / GetRawDualField type 0, engine_interface_fdtd.cpp:263-268,136-141 */
void orc_raw_field(const orc_sim* s, int is_H, const unsigned pos[3], double out[3]);
/* Engine_Interface_FDTD::CalcFastEnergy engine_interface_fdtd.cpp:302-347 */
double orc_energy(const orc_sim* s);
/* ProcessFields::CalcField Common/processfields.cpp:283-409 with the interpolation of
   engine_interface_fdtd.cpp:63-124,150-204; out is {3,nz,ny,nx} x fastest
   (tools/hdf5_file_writer.cpp:286-302). interp: 0 none 1 node 2 cell. */
void orc_dump_field(const orc_sim* s, int is_H, int interp, const unsigned start[3],
                    const unsigned stop[3], float* out);
/* ProcessFieldsFD::Process Common/processfields_fd.cpp:72-107: weight of one sample (lines 84-86)
   and the accumulation field_fd += field_td * weight (lines 88-100), acc interleaved re/im */
void orc_fd_weight(double freq, double T, double dT, unsigned interval, float out[2]);
void orc_fd_accumulate(float* acc, const float* td, size_t n, const float w[2]);
/* ProcessModeMatch::CalcMultipleIntegrals Common/processmodematch.cpp:222-266; out2 = {value, purity ratio} */
void orc_mode_match(const orc_sim* s, int is_H, int ny, const unsigned start[3], const unsigned stop[3],
                    const double* dist0, const double* dist1, double out2[2]);
/* TFSF plane wave (Operator_Ext_TFSF / Engine_Ext_TFSF) on the box of mesh indices start..stop, frequency <= 0
   (phase velocity c0/n); call before orc_build.  Tables: which 0 = voltage, 1 = current, face (n, l), component c */
int orc_set_tfsf(orc_sim* s, const unsigned start[3], const unsigned stop[3], const double prop_dir[3], const double e_amp[3]);
int orc_tfsf_on(const orc_sim* s);
unsigned orc_tfsf_max_delay(const orc_sim* s);
void orc_tfsf_box(const orc_sim* s, unsigned start[3], unsigned stop[3], int active[6]);
unsigned orc_tfsf_face(const orc_sim* s, int which, int n, int l, int c, unsigned* delay, float* delta, float* amp);
/* local absorbing sheets (Operator_Ext_Absorbing_BC / Engine_Ext_Absorbing_BC); x0/x1 mesh indices
   of the sheet, type 1 = MUR_1ST, 2 = MUR_1ST_SA, phase_velocity 0 -> C0; call before orc_build */
int orc_add_absorbing_sheet(orc_sim* s, const unsigned x0[3], const unsigned x1[3], int normal_positive, int type,
                            double phase_velocity);
int orc_abc_count(const orc_sim* s);
void orc_abc_info(const orc_sim* s, int a, int* ny, int* type, int* positive, unsigned x0[3], unsigned x1[3]);
void orc_abc_coeff(const orc_sim* s, int a, float* K1P, float* K1PP, float* K2P, float* K2PP);
/* mesh helpers, operator.cpp:143-206 */
double orc_edge_length(const orc_sim* s, int n, const unsigned pos[3], int dual);
double orc_disc_line(const orc_sim* s, int n, unsigned pos, int dual);


/* ---- sse-compressed + multithreaded engine restatement (fdtd_oracle_sse.c): runs on the
   operator and extensions of a built orc_sim; while it exists the sim's field accessors
   (probes, energy, dumps) read the sse-layout fields. */
void*    orc_sse_create(orc_sim* s, int threads);
void     orc_sse_destroy(void* e);
void     orc_sse_iterate(void* e, unsigned n_ts);
unsigned orc_sse_unique(void* e);   /* number of de-duplicated f4 coefficient tuples */
unsigned orc_sse_num_ts(void* e);
void     orc_sse_get_fields(void* e, float* volt_nijk, float* curr_nijk);

#ifdef __cplusplus
}
#endif
#endif
