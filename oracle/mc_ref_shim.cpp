// TEST INFRASTRUCTURE -- build recipe glue for oracle/_ref/libmc_ref.so.
//
// Compiles the reference's OWN marching-cubes source where it lies
// (/root/reference/third_parties/coslam/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp, with its
// tables.h / sparsegrid3.h / marching_cubes.h; passed on the include path by oracle/build_ref.py) into a shared library with
// a C entry point.  Nothing of the reference is copied: this file only (1) pre-defines the include guard of the reference's
// numpy-bound accessor header pyarraymodule.h and supplies the three names marching_cubes.h needs from it on top of a plain
// double array, and (2) exports marching_cubes() through extern "C".
#include <cassert>      // the headers Python.h pulled in for the reference's file
#include <cmath>
#include <cstddef>
#include <limits>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#define _EXTMODULE_H                       // skip the reference's pyarraymodule.h (needs Python.h / numpy headers)
typedef long npy_intp;
struct PyArrayObject {
  const double* data;
  long dims[3];
};
template <typename T>
T PyArray_SafeGet(const PyArrayObject* a, const npy_intp* c) {
  return static_cast<T>(a->data[(c[0] * a->dims[1] + c[1]) * a->dims[2] + c[2]]);
}

#include "marching_cubes.cpp"              // the reference's file, found through -I

// volume: C-contiguous double [nx,ny,nz] (what mcubes.marching_cubes receives from numpy).  The result arrays are malloc'ed;
// free them with mc_ref_free.  Returns 0.
extern "C" int mc_ref_run(const double* volume, long nx, long ny, long nz, double isovalue, double truncation, double** verts,
                          size_t* n_vert_values, unsigned long** faces, size_t* n_face_values) {
  PyArrayObject arr{volume, {nx, ny, nz}};
  npy_accessor acc(&arr, {nx, ny, nz});
  std::vector<double> v;
  std::vector<unsigned long> f;
  marching_cubes(acc, isovalue, truncation, v, f);
  *n_vert_values = v.size();
  *n_face_values = f.size();
  *verts = static_cast<double*>(std::malloc(sizeof(double) * (v.size() + 1)));
  *faces = static_cast<unsigned long*>(std::malloc(sizeof(unsigned long) * (f.size() + 1)));
  std::memcpy(*verts, v.data(), sizeof(double) * v.size());
  std::memcpy(*faces, f.data(), sizeof(unsigned long) * f.size());
  return 0;
}

extern "C" void mc_ref_free(void* p) { std::free(p); }
