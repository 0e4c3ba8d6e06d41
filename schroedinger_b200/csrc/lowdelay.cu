// lowdelay.cu -- VC-2 / Dirac low-delay slice decoder for sm_100a (second half).
//
// Bit-exact replacement for schro_decoder_decode_lowdelay_transform_data
// (schroedinger/schrolowdelay.c:745-761) and the three functions it dispatches to (:558-742): every
// slice = 7-bit quantiser base index, the length of its luma part, the luma coefficients of the slice's
// codeblock of every subband, then the two chroma planes interleaved (:101-178) -- interleaved
// exp-Golomb signed integers (schroedinger/schrounpack.c:209-246) that read as 1 bits past the end of
// their part (:96-103) -- dequantised with the subband's quantiser (schroedinger/schroutils.c:179-189; the
// s16 "fast" path runs the 16-bit Orc program, schroedinger/schroorc.orc:1204-1217); then DC prediction
// of the LL band (schroedinger/schrodecoder.c:3219-3277).
//
// Slices are independent: ONE THREAD PER SLICE (thousands per picture, times the pictures of a batch),
// so the host uploads the compressed slices -- a few hundred KB per 1080p picture instead of 6 MB of
// coefficients.  A thread keeps 64 bits of its stream in a register, finds a code's end with one
// count-leading-zeros on the follow bits and gathers its data bits with a 4-step bit compress.  DC
// prediction is a wavefront (a sample needs its left, upper and upper-left neighbours): one CTA per
// (picture, component), one thread per LL row, row j one sample behind row j-1.

#include "common.cuh"
#include <climits>

namespace sb2 {

constexpr int LD_MAX_BANDS = 1 + 3 * SB2_DEQUANT_MAX_LEVELS;

struct LowdelayArgs {
  PlaneSet coeffs;
  const uint8_t *data;
  size_t data_pitch;                 // bytes between the pictures' slice buffers
  long long data_bytes;              // bytes of one picture's slices
  int width[3], height[3];
  int depth, nh, nv, n_bytes, remainder, denom, count, orc16;
  int tile_bytes;                    // dynamic shared memory of the slice kernel (0: read the slices from global memory)
};

// quantiser tables (61 entries each, the reference's schro_table_quant / schro_table_offset_1_2) and the
// quantisation matrix: a kernel argument (constant bank), so concurrent calls cannot disturb each other
struct LowdelayTables {
  unsigned quant[61], offset[61];
  int matrix[LD_MAX_BANDS];
};

struct BitReader {
  const uint8_t *data;              // the picture's slice buffer in global memory
  const uint8_t *tile;              // the CTA's slices staged in shared memory: bytes [tile_lo, tile_hi) of the buffer
  long long tile_lo, tile_hi;
  long long pos, end;               // next bit to load into the buffer / first bit past the part
  unsigned long long buf;           // MSB-aligned
  int n;                            // valid bits in buf
  __device__ __forceinline__ void init (const uint8_t *d, const uint8_t *t, long long lo, long long hi, long long p, long long e)
  {
    data = d; tile = t; tile_lo = lo; tile_hi = hi; pos = p; end = e; buf = 0; n = 0;
  }
  __device__ __forceinline__ unsigned byte_at (long long i) const
  {
    // a luma part that runs past its slice may leave the staged range: those bytes come from global memory
    return (i >= tile_lo && i < tile_hi) ? tile[i - tile_lo] : data[i];
  }
  __device__ __forceinline__ void fill ()
  {
    // byte-aligned after the first load; past `end` the stream is all ones (schrounpack.c:96-103)
    while (n <= 56) {
      unsigned b;
      const int sh = (int) (pos & 7);
      if (pos >= end) b = 0xffu;
      else {
        b = byte_at (pos >> 3);
        if (pos - sh + 8 > end) b |= 0xffu >> (int) (end - (pos - sh));
      }
      if (sh) {                      // first, unaligned load: drop the bits before pos
        buf |= (unsigned long long) ((b << sh) & 0xffu) << (56 - n);
        n += 8 - sh; pos += 8 - sh;
      } else {
        buf |= (unsigned long long) b << (56 - n);
        n += 8; pos += 8;
      }
    }
  }
  __device__ __forceinline__ unsigned get (int k)          // k <= 32
  {
    if (k == 0) return 0;
    fill ();
    const unsigned v = (unsigned) (buf >> (64 - k));
    buf <<= k; n -= k;
    return v;
  }
  // schro_unpack_decode_sint: (0 b)* 1 [sign]
  __device__ __forceinline__ int sint ()
  {
    fill ();
    const unsigned x = (unsigned) (buf >> 32);
    const unsigned follow = x & 0xaaaaaaaau;
    if (follow) {
      const int c2 = __clz ((int) follow);                 // = 2 * count, the terminator's position
      const int count = c2 >> 1;
      unsigned y = c2 ? (x >> (32 - c2)) & 0x55555555u : 0u;
      y = (y | (y >> 1)) & 0x33333333u;
      y = (y | (y >> 2)) & 0x0f0f0f0fu;
      y = (y | (y >> 4)) & 0x00ff00ffu;
      y = (y | (y >> 8)) & 0x0000ffffu;
      int v = (int) ((1u << count) - 1u + y);
      int used = c2 + 1;
      if (v) {
        if ((buf >> (63 - used)) & 1) v = -v;
        used++;
      }
      buf <<= used; n -= used;
      return v;
    }
    // sixteen or more data bits: bit by bit, 32-bit wrap-around like the reference's int arithmetic
    unsigned count = 0, value = 0;
    while (!get (1)) { count++; value = (value << 1) | get (1); }
    int v = (int) ((1u << (count & 31)) - 1u + value);
    if (v && get (1)) v = -v;
    return v;
  }
};

__device__ __forceinline__ int ld_dequant (int q, int factor, int offset, int orc16)
{
  if (orc16) {
    const short q16 = (short) q;
    const short s = (short) ((q16 > 0) - (q16 < 0));
    short t = (short) (q16 < 0 ? -q16 : q16);
    t = (short) (t * (short) factor);
    t = (short) (t + (short) (offset + 2));
    t = (short) (t >> 2);
    return (short) (t * s);
  }
  if (q == 0) return 0;
  // int arithmetic of __schro_dequantise, with the wrap-around the reference's build gives it
  if (q < 0) return (int) (0u - (unsigned) (((int) ((0u - (unsigned) q) * (unsigned) factor + (unsigned) offset + 2u)) >> 2));
  return ((int) ((unsigned) q * (unsigned) factor + (unsigned) offset + 2u)) >> 2;
}

constexpr int LD_THREADS = 128;
constexpr int LD_TILE_BYTES = 48 * 1024;       // staged when the CTA's slices fit (128 slices of up to 384 bytes)

template <typename T>
__global__ void __launch_bounds__ (LD_THREADS)
lowdelay_slice_kernel (const LowdelayArgs A, const LowdelayTables c_ld)
{
  extern __shared__ __align__ (16) uint8_t tile[];
  const int per_pic = A.nh * A.nv;
  // a CTA's threads take consecutive slices of ONE picture, so their bytes are one contiguous range
  const int ctas_per_pic = (per_pic + LD_THREADS - 1) / LD_THREADS;
  const int pic = blockIdx.x / ctas_per_pic, k0 = (blockIdx.x - pic * ctas_per_pic) * LD_THREADS;
  const int k = k0 + threadIdx.x;
  const uint8_t *data = A.data + (size_t) pic * A.data_pitch;
  // slice k starts after k * n_bytes + floor (k * remainder / denom) bytes (the accumulator of :620-632)
  auto slice_off = [&] (int kk) { return (long long) kk * A.n_bytes + ((long long) kk * A.remainder) / A.denom; };
  const int k1 = min (k0 + LD_THREADS, per_pic);
  const long long lo = slice_off (k0), hi = slice_off (k1);
  long long tile_lo = 0, tile_hi = 0;
  if (A.tile_bytes >= hi - lo) {
    // coalesced copy of the CTA's slices: the bit readers then run out of shared memory, with no global-memory
    // latency inside their serial chains
    const long long a0 = lo & ~15LL;
    for (long long i = a0 + 16LL * threadIdx.x; i < hi; i += 16LL * LD_THREADS) {
      if (i + 16 <= A.data_bytes && (((size_t) data) & 15) == 0)
        *reinterpret_cast<uint4 *> (tile + (i - a0)) = *reinterpret_cast<const uint4 *> (data + i);
      else
        for (int b = 0; b < 16; b++) if (i + b < A.data_bytes) tile[i - a0 + b] = data[i + b];
    }
    tile_lo = a0; tile_hi = min (hi, A.data_bytes);
  }
  __syncthreads ();
  if (k >= per_pic) return;
  const int sy = k / A.nh, sx = k - sy * A.nh;
  const long long off = slice_off (k);
  const int slice_bytes = (int) (slice_off (k + 1) - off);
  BitReader yb, uvb;
  yb.init (data, tile, tile_lo, tile_hi, 8 * off, 8 * (off + slice_bytes));
  const int base_index = (int) yb.get (7);
  const int lenbits = 32 - __clz (8 * (A.orc16 ? A.n_bytes : slice_bytes));        // ilog2up (:87-97)
  const long long y_length = (long long) yb.get (lenbits);
  const long long y_start = 8 * off + 7 + lenbits;
  // the luma part ends where its declared length says (even past the slice, schrounpack.c:49-60), only the
  // end of the picture's buffer stops it; the chroma part runs from there to the end of the slice
  yb.init (data, tile, tile_lo, tile_hi, y_start, min (y_start + y_length, 8 * A.data_bytes));
  uvb.init (data, tile, tile_lo, tile_hi, y_start + y_length, 8 * (off + slice_bytes));
  const int nbands = 1 + 3 * A.depth;
#pragma unroll 1
  for (int c = 0; c < 2; c++) {
    T *p0 = reinterpret_cast<T *> (plane_ptr (A.coeffs, pic, c ? 1 : 0));
    T *p1 = reinterpret_cast<T *> (plane_ptr (A.coeffs, pic, 2));
    const int stride = A.coeffs.stride[c ? 1 : 0] / (int) sizeof (T);
    const int stride2 = A.coeffs.stride[2] / (int) sizeof (T);
    const int W = A.width[c], H = A.height[c];
#pragma unroll 1
    for (int i = 0; i < nbands; i++) {
      const int qi = min (max (base_index - c_ld.matrix[i], 0), 60);
      const int factor = (int) c_ld.quant[qi], qoff = (int) c_ld.offset[qi];
      const int level = i == 0 ? 0 : (i - 1) / 3, orient = i == 0 ? 0 : (i - 1) % 3 + 1;
      const int shift = A.depth - level;
      const int bw = W >> shift, bh = H >> shift;
      const int x0 = bw * sx / A.nh, x1 = bw * (sx + 1) / A.nh;
      const int y0 = bh * sy / A.nv, y1 = bh * (sy + 1) / A.nv;
      const long long first = ((orient & 2) ? ((long long) stride << shift) >> 1 : 0) + ((orient & 1) ? bw : 0);
      const long long first2 = ((orient & 2) ? ((long long) stride2 << shift) >> 1 : 0) + ((orient & 1) ? bw : 0);
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          if (c == 0) {
            p0[first + ((long long) y * stride << shift) + x] = (T) ld_dequant (yb.sint (), factor, qoff, A.orc16);
          } else {
            const int u = ld_dequant (uvb.sint (), factor, qoff, A.orc16);
            const int v = ld_dequant (uvb.sint (), factor, qoff, A.orc16);
            p0[first + ((long long) y * stride << shift) + x] = (T) u;
            p1[first2 + ((long long) y * stride2 << shift) + x] = (T) v;
          }
        }
    }
  }
}

// schro_decoder_subband_dc_predict / _s32: in place on the LL band
template <typename T>
__global__ void __launch_bounds__ (1024)
lowdelay_dc_kernel (const LowdelayArgs A)
{
  __shared__ int ring[3][1024];
  const int comp = blockIdx.x % 3, pic = blockIdx.x / 3;
  const int w = A.width[comp ? 1 : 0] >> A.depth, h = A.height[comp ? 1 : 0] >> A.depth;
  const int j = threadIdx.x;
  T *row = reinterpret_cast<T *> (plane_ptr (A.coeffs, pic, comp)) + ((long long) j * (A.coeffs.stride[comp] / (int) sizeof (T)) << A.depth);
  int left = 0;
  for (int s = 0; s < w + h - 1; s++) {
    const int i = s - j;
    if (j < h && i >= 0 && i < w) {
      int v = (int) row[i];
      if (j == 0) { if (i > 0) v += left; }
      else if (i == 0) v += ring[(s + 2) % 3][j - 1];
      else {
        const int a = (int) ((unsigned) left + (unsigned) ring[(s + 2) % 3][j - 1] + (unsigned) ring[(s + 1) % 3][j - 1] + 1u);
        // schro_divide3 for s16 (schroutils.h:64), schro_divide (a, 3) for s32 (:63); int arithmetic wraps
        const int pred = sizeof (T) == 2 ? (int) ((unsigned) a * 21845u + 10922u) >> 16 : (a < 0 ? (int) ((unsigned) a - 2u) / 3 : a / 3);
        v = (int) ((unsigned) v + (unsigned) pred);
      }
      v = (int) (T) v;
      row[i] = (T) v;
      left = v;
      ring[s % 3][j] = v;
    }
    __syncthreads ();
  }
}

}  // namespace sb2

using namespace sb2;

// tests run the slice kernel both ways: slices staged in shared memory (default) and read from global memory
static int g_lowdelay_unstaged = 0;
extern "C" void sb2_lowdelay_force_unstaged (int on) { g_lowdelay_unstaged = on ? 1 : 0; }

extern "C" int
sb2_lowdelay_decode (const sb2_lowdelay_params *p, const uint8_t *slices, size_t picture_bytes, size_t picture_pitch,
    const sb2_slab *coeffs, int is_s32, void *stream)
{
  if (!p || !slices || !coeffs || !coeffs->base) return set_error (SB2_ERR_ARG, "sb2_lowdelay_decode: null argument");
  if (coeffs->ncomp != 3 || coeffs->count < 1) return set_error (SB2_ERR_ARG, "sb2_lowdelay_decode: needs three-component coefficient frames");
  if (p->transform_depth < 1 || p->transform_depth > SB2_DEQUANT_MAX_LEVELS || p->n_horiz_slices < 1 || p->n_vert_slices < 1 ||
      p->slice_bytes_denom < 1 || p->slice_bytes_num < p->slice_bytes_denom)
    return set_error (SB2_ERR_ARG, "sb2_lowdelay_decode: bad parameters");
  if (coeffs->width[1] != coeffs->width[2] || coeffs->height[1] != coeffs->height[2])
    return set_error (SB2_ERR_ARG, "sb2_lowdelay_decode: the chroma planes differ in size");
  const int bpp = is_s32 ? 4 : 2;
  for (int c = 0; c < 3; c++) {
    if ((coeffs->stride[c] % bpp) || (coeffs->offset[c] % bpp)) return set_error (SB2_ERR_ARG, "sb2_lowdelay_decode: unaligned plane");
    if ((coeffs->height[c] >> p->transform_depth) > 1024)
      return set_error (SB2_ERR_UNSUPPORTED, "sb2_lowdelay_decode: LL band taller than 1024 rows");
  }
  const long long n_bytes = p->slice_bytes_num / p->slice_bytes_denom, rem = p->slice_bytes_num % p->slice_bytes_denom;
  const long long nslices = (long long) p->n_horiz_slices * p->n_vert_slices;
  const long long total = nslices * n_bytes + (nslices * rem) / p->slice_bytes_denom;
  if ((long long) picture_bytes < total)
    return set_error (SB2_ERR_ARG, "sb2_lowdelay_decode: %zu bytes of slices per picture, the parameters need %lld", picture_bytes, total);
  LowdelayTables t;
  for (int i = 0; i < 61; i++) { t.quant[i] = p->table_quant[i]; t.offset[i] = p->table_offset[i]; }
  for (int i = 0; i < LD_MAX_BANDS; i++) t.matrix[i] = i < 1 + 3 * p->transform_depth ? p->quant_matrix[i] : 0;
  cudaStream_t st = as_stream (stream);
  LowdelayArgs A;
  A.coeffs = planeset_from_slab (coeffs);
  A.data = slices;
  A.data_pitch = picture_pitch;
  A.data_bytes = (long long) picture_bytes;
  A.width[0] = coeffs->width[0]; A.height[0] = coeffs->height[0];
  A.width[1] = A.width[2] = coeffs->width[1]; A.height[1] = A.height[2] = coeffs->height[1];
  A.depth = p->transform_depth;
  A.nh = p->n_horiz_slices;
  A.nv = p->n_vert_slices;
  A.n_bytes = (int) n_bytes;
  A.remainder = (int) rem;
  A.denom = p->slice_bytes_denom;
  A.count = coeffs->count;
  // the reference's dispatcher (schrolowdelay.c:745-761): s32 frames and slice grids that do not divide the
  // chroma LL band take the plain-C dequantiser, the rest the 16-bit Orc program
  A.orc16 = !is_s32 && ((coeffs->width[1] >> p->transform_depth) % p->n_horiz_slices) == 0 &&
      ((coeffs->height[1] >> p->transform_depth) % p->n_vert_slices) == 0;
  double coef = 0;
  for (int c = 0; c < 3; c++) coef += (double) coeffs->width[c] * coeffs->height[c];
  {
    LaunchScope scope ("lowdelay_slices", ((double) total + coef * bpp) * coeffs->count, st);
    const long long ctas_per_pic = (nslices + LD_THREADS - 1) / LD_THREADS;
    const unsigned ctas = (unsigned) (ctas_per_pic * coeffs->count);
    // bytes of LD_THREADS consecutive slices (+ one for the fractional sizes, + alignment slack)
    const long long need = (n_bytes + 1) * LD_THREADS + 32;
    A.tile_bytes = (need <= LD_TILE_BYTES && !g_lowdelay_unstaged) ? (int) need : 0;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute (lowdelay_slice_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, LD_TILE_BYTES);
      cudaFuncSetAttribute (lowdelay_slice_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, LD_TILE_BYTES);
      attr_set = true;
    }
    if (is_s32) lowdelay_slice_kernel<int32_t><<<ctas, LD_THREADS, A.tile_bytes, st>>> (A, t);
    else lowdelay_slice_kernel<int16_t><<<ctas, LD_THREADS, A.tile_bytes, st>>> (A, t);
  }
  {
    LaunchScope scope ("lowdelay_dc_predict", 2.0 * coef / (1 << (2 * p->transform_depth)) * bpp * coeffs->count, st);
    const int threads = min (1024, ((coeffs->height[0] >> p->transform_depth) + 31) & ~31);
    if (is_s32) lowdelay_dc_kernel<int32_t><<<3 * coeffs->count, threads, 0, st>>> (A);
    else lowdelay_dc_kernel<int16_t><<<3 * coeffs->count, threads, 0, st>>> (A);
  }
  return check_cuda (cudaGetLastError (), "lowdelay kernels launch");
}
