// Host-sequenced implementations behind mdkm_ground_level / mdkm_kmeans_plusplus.
#pragma once
#include "extras.cuh"

namespace mdkm {

inline int ground_level_impl(cudaStream_t, float*, long long, float*, int, double*, double*, int*) {
  return -1;  // TODO(next row f2)
}

inline int kmeanspp_impl(cudaStream_t, int, const float*, long long, FrameF, int, long long,
                         const double*, int, double*, long long*, int*) {
  return -1;  // TODO(next row f1)
}

}  // namespace mdkm
