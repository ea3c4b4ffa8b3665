// synthetic_operator.h -- host-side operator build of Operator_CUDA for box geometries.
//
// In an openEMS build, Operator_CUDA derives from Operator_Multithread and reuses the
// reference's CalcECOperator; it only re-keys and uploads the result.  That path
// needs CSXCAD and 96 B/cell of host RAM (103 GB at 1024^3, SURVEY 7 "hard parts"), so this
// library also carries its own builder: the same formulas (FDTD/operator.cpp:956-984,
// 1099-1187, 1347-1444, 1956-2030; extensions/operator_ext_upml.cpp:269-445,
// operator_ext_mur_abc.cpp:107-186, operator_ext_excitation.cpp:105-297), evaluated plane by
// plane and emitted directly in the compressed device format (table + per-cell index), with
// planes of equal z-signature computed once.  Geometry is a list of axis-aligned boxes
// (stand-in for CSXCAD primitives); excitation is a Gauss pulse or sinus.
#pragma once
#include "../../../include/openems_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
/* The library is built with -fvisibility=hidden: only the C entry points declared here are exported.  Its internal
   C++ classes (one of them is called Engine, like FDTD/engine.h's) must never be visible to, or be interposed by,
   the host application's symbols. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct oems_synth oems_synth;

oems_synth* oems_synth_create(unsigned nx, unsigned ny, unsigned nz, const double* x, const double* y,
                              const double* z, double grid_delta);
void oems_synth_destroy(oems_synth* s);
/* bc: 0 PEC, 1 PMC, 2 MUR, 3 PML (openems.cpp:383-409); pml_size lines per face */
void oems_synth_set_bc(oems_synth* s, const int bc[6], const unsigned pml_size[6]);
void oems_synth_set_background(oems_synth* s, double epsR, double mueR, double kappa, double sigma);
void oems_synth_set_timestep(oems_synth* s, double forced_dT, double factor);
int  oems_synth_add_material(oems_synth* s, int prio, const double start[3], const double stop[3],
                             double epsR, double mueR, double kappa, double sigma);
int  oems_synth_add_metal(oems_synth* s, int prio, const double start[3], const double stop[3]);
/* Drude/Lorentz material box; arrays of length `order` (may be NULL = 0) */
int  oems_synth_add_lorentz(oems_synth* s, int prio, const double start[3], const double stop[3],
                            double epsR, double mueR, double kappa, double sigma, int order,
                            const double* eps_fp, const double* eps_tau, const double* eps_flor,
                            const double* mue_fp, const double* mue_tau, const double* mue_flor);
int  oems_synth_add_excitation(oems_synth* s, int prio, const double start[3], const double stop[3],
                               int exc_type, const double vec[3], double delay_s);
void oems_synth_set_excite_gauss(oems_synth* s, double f0, double fc);
void oems_synth_set_excite_sinus(oems_synth* s, double f0);
/* builds timestep, compressed operator, extension data; 0 on success */
/* z-slab restricted build for one-process-per-GPU runs: only the xy planes rank-owned [z_begin, z_end) (+ one ghost
   plane per side) are built.  The ranks first agree on the timestep: each calls oems_synth_local_timestep (minimum of
   Operator::CalcTimestep over ITS planes), they MIN-reduce the values and hand the result to oems_synth_set_timestep. */
int  oems_synth_set_slab(oems_synth* s, unsigned z_begin, unsigned z_end);
int  oems_synth_local_timestep(oems_synth* s, double* dT_out);
int  oems_synth_build(oems_synth* s, unsigned max_ts);
const char* oems_synth_last_error(const oems_synth* s);

double   oems_synth_dT(const oems_synth* s);
unsigned oems_synth_nyquist(const oems_synth* s);
unsigned oems_synth_n_unique(const oems_synth* s);
int      oems_synth_index_bytes(const oems_synth* s);
const oems_coeff_entry* oems_synth_table(const oems_synth* s);
const void* oems_synth_index(const oems_synth* s); /* [nz][ny][nx] */
/* the same index as unique xy planes [unique_planes][ny][nx] + one plane id per z (what oems_synth_upload sends) */
const unsigned* oems_synth_plane_of_z(const oems_synth* s);
const void* oems_synth_plane_data(const oems_synth* s);
unsigned oems_synth_unique_planes(const oems_synth* s);
unsigned oems_synth_signal_length(const oems_synth* s);
const float* oems_synth_signal(const oems_synth* s, int is_curr);
unsigned oems_synth_exc_count(const oems_synth* s, int is_curr);
void oems_synth_exc_get(const oems_synth* s, int is_curr, unsigned* idx3, unsigned* dir, float* amp, unsigned* delay);
int  oems_synth_mur_count(const oems_synth* s);
const float* oems_synth_mur_coeff(const oems_synth* s, int m, int which, int* ny, unsigned* line, unsigned* shift,
                                  unsigned nlines[2], unsigned* start_ts);
int  oems_synth_upml_count(const oems_synth* s);
void oems_synth_upml_box(const oems_synth* s, int b, unsigned start[3], unsigned nlines[3]);
int  oems_synth_lorentz_order(const oems_synth* s);
unsigned oems_synth_lorentz_count(const oems_synth* s, int o);

/* uploads everything into a created (not yet finalized) engine and finalizes it */
int oems_synth_upload(const oems_synth* s, oems_cuda_engine* eng);
/* page-lock the index buffer (needs a CUDA device); engine uploads then run at the PCIe rate */
int oems_synth_pin(oems_synth* s);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
