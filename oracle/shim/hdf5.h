/* Minimal stand-in for <hdf5.h>: the reference's SquiggleRead header includes
 * its fast5 I/O header, which only needs these three typedef names to parse.
 * No HDF5 function is ever called by the oracle build (test infrastructure). */
#ifndef ORACLE_SHIM_HDF5_H
#define ORACLE_SHIM_HDF5_H
#include <stdint.h>
typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#endif
