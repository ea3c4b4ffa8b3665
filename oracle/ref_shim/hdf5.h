/* TEST INFRASTRUCTURE shim: the two HDF5 typedefs tools/hdf5_file_writer.h:39-58 names in its declarations.
 * tools/hdf5_file_writer.cpp is not compiled; oracle/ref_glue.cpp implements HDF5_File_Writer as an
 * in-memory recorder so that the reference's ProcessFields* dump path can be observed. */
#pragma once
#include <stdint.h>
typedef int64_t hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
