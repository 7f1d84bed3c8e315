// Device-side interface of the fused energy-balance kernels (kernels.cu), used by api.cu.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace enrgy {

// insolation source of the fused kernel (template parameter)
constexpr int kInsolStreamed = 0;  // per-step kWh m-2 raster streamed from HBM
constexpr int kInsolComputed = 1;  // terrain normal . sun vector per sub-step, no shadows
constexpr int kInsolShadow = 2;    // ... with the ray-marched sunlit mask
constexpr int kInsolShadowKeys = 3;  // ... the same, sampling the integer copy of the DEM (no negative elevations)
__host__ __device__ constexpr bool insol_shadow(int insol) { return insol >= kInsolShadow; }

// Max pyramid of the DEM: level l holds the max of the valid cells of every (16 << l)^2 block as
// [(nbr + 2)][(nbc + 2)] floats with one ring of -inf blocks, at offset off[l] of one buffer.
constexpr int kMaxPyramidLevels = 12;   // 16 << 11 = 32768 >= the largest raster edge
struct MaxPyramid {
  int levels;
  int off[kMaxPyramidLevels], nbr[kMaxPyramidLevels], nbc[kMaxPyramidLevels];
};

template <typename R>
struct KernelArgs {
  // geometry (padded device rasters; pitch is a multiple of kTileW, padding cells hold NaN)
  int rows_full, cols, pitch;     // full DEM
  int band_row0, band_rows;       // this handle's row band (rasters below are band-local)
  int rows_pad_full;              // padded row count of the full DEM buffer
  const float* dem;               // full DEM (replicated for shading), pointing at cell (0, 0) of a
                                  // buffer with a NaN apron of kDemApron cells on every side
  int dem_pitch;                  // row stride of the DEM buffer
  const int* dem_keys;            // same layout: -(bit pattern) of every valid cell, +1 for NaN (dem_key_kernel)
  const float* blockmax;          // max pyramid of the DEM (shading early exit), see MaxPyramid
  MaxPyramid pyramid;
  const float* gstep;             // 8 step-rise pyramids (one per ray octant, same layout), see gstep_kernel
  int pyr_stride;                 // floats per pyramid
  const R* nx;                    // [band_rows_pad][pitch] terrain normal (computed insolation)
  const R* ny;
  const R* nz;
  const float* albedo;            // [n_maps][band_rows_pad][pitch]
  size_t map_stride;              // elements between consecutive albedo maps
  int albedo_const;
  R albedo_ice, albedo_snow, max_ice_albedo;
  R albedo_offset;                // ensemble member: added to the map values, clipped to [0.001, 1]
  R elev_aws;
  R zmax;                         // max of the valid DEM, as float exactly
  // state, band-local [band_rows_pad][pitch]
  R* swe;
  R* total_snow;
  R* total_ice;
  // sub-surface model: boundary temperatures [layers + 1][band_rows_pad][pitch] (deg C)
  R* layer_t;
  size_t layer_stride;
  MsmParams<R> msm;
  // streamed insolation [n_steps_resident][band_rows_pad][pitch], step index relative to pot_t0
  const float* pot;
  size_t pot_stride;
  int pot_t0;
  // tables
  const StepRec<R>* steps;        // [n_steps]
  const SubRec<R>* subs;          // [n_subs]
  const ShadeRec* shades;         // [n_subs]
  const TimeBlock* blocks;        // [n_blocks] covering at least [t0, t1)
  int cap_steps, cap_subs;        // capacities the time blocks were cut for (smem staging buffers)
  int block_begin, block_end;     // blocks to process
  int t0, t1;                     // steps to process (clip of the first/last block)
  // work list
  const int2* tiles;              // active tiles (tile row, tile col) of the band
  int n_tiles;
  // statistics: per-CTA partial sums [gridDim.x][t1 - t0][kStatsP] float64
  R* partials;                    // [gridDim.x][t1 - t0][kStatsK (+ kStatsM with the sub-surface model)] in R
  // dump mode: [t1 - t0][ENRGY_D_COUNT][band_rows_pad][pitch] R (may be null)
  R* dump;
  size_t dump_field_stride;
  // shade-mask dump: [n_sub][band_rows][words] bit masks of step t0 (may be null)
  unsigned* mask_out;
  int mask_words;
};

struct FinalizeArgs {
  const void* partials;     // [n_ctas][n_steps][row], row = kStatsK (+ kStatsM with msm); float if f32_mode else double
  int msm;                  // sub-surface model on: lwu and g are summed per cell
  int n_ctas, n_steps, t0;
  double n_valid;           // valid cells of the band
  double mom[5];            // sum over the band's glacier cells of (dem - elev_aws)^k, k = 0..4
  int f32_mode;             // round the per-step constants the way the float32 kernel saw them
  const StepRec<double>* steps64;  // master copy of the per-step records (float64)
  double* stats;            // [n_steps][ENRGY_S_COUNT]
  // first-row quirk of the reference (model.py:248-252): SWE statistics of step 0 run over every
  // non-NaN cell of the INITIAL raster, including off-glacier cells
  int override_first;
  double swe0_sum, swe0_nsnow, swe0_nvalid;
};

// launch helpers (defined in kernels.cu); all asynchronous on `stream`
template <typename R>
cudaError_t launch_terrain(const float* dem, int dem_pitch, int rows_full, int cols, int pitch,
                           int band_row0, int band_rows_pad, double cell, R* nx, R* ny, R* nz,
                           cudaStream_t stream);
cudaError_t launch_blockmax(const float* dem, int dem_pitch, int rows_full, int cols, const MaxPyramid& py,
                            float* buffer, cudaStream_t stream);
// integer copy of the DEM buffer for the shading samples + min of the valid cells (as float bits key)
cudaError_t launch_dem_keys(const float* dem_buf, int* key_buf, size_t n, int* min_key /*device, preset INT_MAX*/,
                            cudaStream_t stream);
// step-rise pyramids of the 8 ray octants: buffer[oct * stride + pyramid layout], stride = floats per pyramid
cudaError_t launch_gstep(const float* dem, int dem_pitch, int rows_full, int cols, const MaxPyramid& py, int stride,
                         float* buffer, cudaStream_t stream);
cudaError_t launch_tile_scan(const float* dem, int pitch, int band_row0, int band_rows, int cols,
                             int tile_h, int tile_w, int tiles_r, int tiles_c, int* counts,
                             cudaStream_t stream);
cudaError_t launch_mask_check(const float* dem, int dem_pitch, const float* other, int pitch,
                              int band_row0, int band_rows, int cols,
                              unsigned long long* counters /*[2]*/, cudaStream_t stream);
template <typename R>
cudaError_t launch_moments(const float* dem, int dem_pitch, int band_row0, int band_rows, int cols, double elev,
                           double* block_out /*[blocks][5]*/, int blocks, cudaStream_t stream);
cudaError_t launch_swe0_stats(const float* swe, int pitch, int band_rows, int cols,
                              double* block_out /*[blocks][3]*/, int blocks, cudaStream_t stream);
template <typename R>
cudaError_t launch_pad_convert(const float* src_pitched, R* dst, size_t n, cudaStream_t stream);
template <typename R>
cudaError_t launch_unpad_state(const R* src, int pitch, int rows, int cols, int dtype, void* dst,
                               cudaStream_t stream);

template <typename R>
cudaError_t launch_nan_offglacier(const float* dem, int dem_pitch, int pitch, int band_row0,
                                  int band_rows, int cols, R* swe, R* tsn, R* tic, cudaStream_t stream);

struct LaunchInfo {
  int regs, smem_bytes, ctas_per_sm, grid, cells_per_thread;
};
template <typename R>
cudaError_t launch_energy_balance(const KernelArgs<R>& a, const void* reserved, int insol, bool dump,
                                  int sm_count, int forced_grid, LaunchInfo* info, cudaStream_t stream);
template <typename R>
void energy_balance_tile(bool msm, int insol, int* tile_h, int* tile_w);
template <typename R>
cudaError_t energy_balance_grid(int insol, bool msm, bool dump, int sm_count, int cap_steps, int cap_subs,
                                LaunchInfo* info);
// initial boundary temperatures: min(0, t_point[l] + (dem - elev) * -0.006), model.py:133-143
template <typename R>
cudaError_t launch_msm_init(const float* dem, int dem_pitch, int pitch, int band_row0, int band_rows_pad,
                            int n_bounds, const double* t_point /*host*/, double elev, R* layer_t,
                            size_t layer_stride, cudaStream_t stream);

cudaError_t launch_finalize(const FinalizeArgs& f, cudaStream_t stream);
cudaError_t launch_microbench(int kind, int sm_count, int iters, void* scratch, double* ops_per_launch,
                              cudaStream_t stream);

}  // namespace enrgy
