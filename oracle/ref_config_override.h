/* Force-included (nvcc -include) in front of the reference's .cu files by oracle/build_ref.py.
 * TEST INFRASTRUCTURE ONLY — this is how the 16-d language-feature variant of the UNMODIFIED
 * reference is built: cuda_rasterizer/config.h:15-20 holds bare #defines, so we pre-define its
 * include guard and provide the six constants here, with LSX_REF_F (from -D) as the feature width.
 * CUB is pulled in first because its templates use the identifier NUM_CHANNELS, which must not be
 * a macro yet when those headers are parsed. */
#pragma once
#include <cstdint>
#include <cfloat>
#if defined(__CUDACC__)
#include <cub/cub.cuh>
#include <cub/device/device_radix_sort.cuh>
#endif
#ifndef LSX_REF_F
#error "pass -DLSX_REF_F=<language feature width>"
#endif
#define CUDA_RASTERIZER_CONFIG_H_INCLUDED
#define NUM_CHANNELS 3
#define NUM_CHANNELS_language_feature LSX_REF_F
#define NUM_CHANNELS_instance_feature 3
#define NUM_ALL_MAP 5
#define BLOCK_X 16
#define BLOCK_Y 16
