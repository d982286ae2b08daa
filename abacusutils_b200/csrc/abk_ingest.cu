// Device-side particle ingest: decode Abacus RVint and pack9 particle records that were
// copied to the GPU still packed, so positions never exist on the host.
//   RVint  abacusnbody/data/bitpacked.py:29-120   one thread per int32, HBM-bound (4 B in, 4-16 B out)
//   PIDs   abacusnbody/data/bitpacked.py:123-311  one thread per packed 64-bit aux word
//   pack9  abacusnbody/data/pack9.py:16-123       9-byte records; a record whose first byte is 0xFF is a cell
//          header that sets the origin/scales of the particle records after it.  The reference walks the
//          stream serially; here the "last header before me" dependency is a prefix count of header flags:
//          count per 256-record block -> scan over blocks -> table of decoded headers -> decode particles,
//          each particle reading header number (#headers before it) - 1 and writing row (index - #headers).
// The per-record arithmetic lives in abk_ingest.cuh (shared with the host check that pins it to the reference).
#include "abk_common.cuh"
#include "abk_ingest.cuh"

namespace {

constexpr int P9_BLOCK = 256;                 // records per block, one per thread
constexpr int P9_BYTES = P9_BLOCK * 9;        // 2304 bytes, a multiple of 4
constexpr int P9_WORDS = P9_BYTES / 4;

template <typename T>
__global__ void __launch_bounds__(256) rvint_kernel(const int32_t *__restrict__ in, int64_t n3, double posscale,
                                                    T *__restrict__ pos, T *__restrict__ vel)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += stride) {
        const int32_t v = in[i];
        if (pos) pos[i] = abk_rvint_pos<T>(v, posscale);
        if (vel) vel[i] = abk_rvint_vel<T>(v);
    }
}

// stage this block's 256 records in shared memory with coalesced 4-byte loads (data is 4-byte aligned and
// the block offset 2304*b keeps that alignment); bytes past the end of the stream read as 0
__device__ __forceinline__ void p9_stage(const uint8_t *__restrict__ data, int64_t nrec, uint8_t *s_bytes)
{
    const int64_t off = (int64_t)blockIdx.x * P9_BYTES;
    const int64_t total = nrec * 9;
    const int64_t left = total - off;
    const int nbytes = left < P9_BYTES ? (int)left : P9_BYTES;
    const int nfull = nbytes >> 2;
    const uint32_t *src = (const uint32_t *)(data + off);
    uint32_t *dst = (uint32_t *)s_bytes;
    for (int t = threadIdx.x; t < P9_WORDS; t += blockDim.x) dst[t] = t < nfull ? src[t] : 0u;
    __syncthreads();
    if (threadIdx.x < (nbytes & 3)) s_bytes[nfull * 4 + threadIdx.x] = data[off + nfull * 4 + threadIdx.x];
    __syncthreads();
}

// number of header records before this thread's record inside the block; *is_hdr = this record is a header
__device__ __forceinline__ unsigned p9_block_prefix(bool is_hdr, unsigned *s_warp)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(full, is_hdr);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    unsigned before = __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; w++) before += s_warp[w];
    return before;
}

__global__ void __launch_bounds__(P9_BLOCK) pack9_count_kernel(const uint8_t *__restrict__ data, int64_t nrec,
                                                               uint32_t *__restrict__ blk_hdr)
{
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * P9_BLOCK + threadIdx.x;
    const bool is_hdr = i < nrec && data[i * 9] == 0xFF;
    const unsigned bal = __ballot_sync(0xffffffffu, is_hdr);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&s_cnt, (unsigned)__popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) blk_hdr[blockIdx.x] = s_cnt;
}

template <typename T>
__global__ void __launch_bounds__(P9_BLOCK) pack9_headers_kernel(const uint8_t *__restrict__ data, int64_t nrec,
                                                                 const uint32_t *__restrict__ blk_incl, T boxsize, T velz,
                                                                 abk_pack9_cell<T> *__restrict__ hdr_tab)
{
    __shared__ __align__(16) uint8_t s_bytes[P9_BYTES];
    __shared__ unsigned s_warp[P9_BLOCK / 32];
    p9_stage(data, nrec, s_bytes);
    const int64_t i = (int64_t)blockIdx.x * P9_BLOCK + threadIdx.x;
    const uint8_t *c = s_bytes + threadIdx.x * 9;
    const bool is_hdr = i < nrec && abk_pack9_is_header(c);
    const unsigned before = p9_block_prefix(is_hdr, s_warp);
    if (is_hdr) {
        const int64_t h = (int64_t)(blockIdx.x ? blk_incl[blockIdx.x - 1] : 0u) + before;
        int s[6];
        abk_pack9_expand(c, s);
        hdr_tab[h] = abk_pack9_header<T>(s, boxsize, velz);
    }
}

template <typename T>
__global__ void __launch_bounds__(P9_BLOCK) pack9_decode_kernel(const uint8_t *__restrict__ data, int64_t nrec,
                                                                const uint32_t *__restrict__ blk_incl,
                                                                const abk_pack9_cell<T> *__restrict__ hdr_tab,
                                                                T *__restrict__ pos, T *__restrict__ vel)
{
    __shared__ __align__(16) uint8_t s_bytes[P9_BYTES];
    __shared__ unsigned s_warp[P9_BLOCK / 32];
    p9_stage(data, nrec, s_bytes);
    const int64_t i = (int64_t)blockIdx.x * P9_BLOCK + threadIdx.x;
    const uint8_t *c = s_bytes + threadIdx.x * 9;
    const bool valid = i < nrec;
    const bool is_hdr = valid && abk_pack9_is_header(c);
    const unsigned before = p9_block_prefix(is_hdr, s_warp);
    if (!valid || is_hdr) return;
    const int64_t nh = (int64_t)(blockIdx.x ? blk_incl[blockIdx.x - 1] : 0u) + before;   // headers before record i
    abk_pack9_cell<T> h;
    if (nh > 0) {
        h = hdr_tab[nh - 1];
    } else {   // particle before any header: the reference's header state is still NaN (pack9.py:67-72)
        const T nan = (T)__longlong_as_double(0x7ff8000000000000LL);
        h.pscale = h.cellx = h.celly = h.cellz = h.vscale = nan;
    }
    int s[6];
    abk_pack9_expand(c, s);
    T p[3], v[3];
    abk_pack9_particle<T>(s, h, p, v);
    const int64_t w = i - nh;
    if (pos) { pos[3 * w] = p[0]; pos[3 * w + 1] = p[1]; pos[3 * w + 2] = p[2]; }
    if (vel) { vel[3 * w] = v[0]; vel[3 * w + 1] = v[1]; vel[3 * w + 2] = v[2]; }
}

template <typename T>
__global__ void __launch_bounds__(256) pids_kernel(const uint64_t *__restrict__ packed, int64_t n, T inv_ppd, T half,
                                                   int64_t *__restrict__ pid, T *__restrict__ lagr_pos,
                                                   int16_t *__restrict__ lagr_idx, uint8_t *__restrict__ tagged,
                                                   T *__restrict__ density)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t p = packed[i];
        if (pid) pid[i] = abk_pid_pid(p);
        if (lagr_pos) {
            T v[3];
            abk_pid_lagr_pos<T>(p, inv_ppd, half, v);
            lagr_pos[3 * i] = v[0]; lagr_pos[3 * i + 1] = v[1]; lagr_pos[3 * i + 2] = v[2];
        }
        if (lagr_idx) {
            int16_t v[3];
            abk_pid_lagr_idx(p, v);
            lagr_idx[3 * i] = v[0]; lagr_idx[3 * i + 1] = v[1]; lagr_idx[3 * i + 2] = v[2];
        }
        if (tagged) tagged[i] = abk_pid_tagged(p);
        if (density) density[i] = abk_pid_density<T>(p);
    }
}

int64_t p9_blocks(int64_t nrec) { return (nrec + P9_BLOCK - 1) / P9_BLOCK; }

}  // namespace

extern "C" int abk_unpack_rvint(abk_ctx *ctx, const int32_t *intdata, int64_t N, double boxsize, void *posout, void *velout,
                                int out_f64)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && N >= 0 && (N == 0 || intdata), "abk_unpack_rvint: bad arguments");
    if (N == 0 || (!posout && !velout)) return ABK_OK;
    const int64_t n3 = 3 * N;
    int64_t blocks = (n3 + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    const double posscale = boxsize / 1e6;   // bitpacked.py:105
    if (out_f64)
        ABK_LAUNCH(ctx, ABK_K_MISC, rvint_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                                        intdata, n3, posscale, (double *)posout, (double *)velout));
    else
        ABK_LAUNCH(ctx, ABK_K_MISC, rvint_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                                        intdata, n3, posscale, (float *)posout, (float *)velout));
    return ABK_OK;
}

extern "C" int abk_pack9_scratch_bytes(int64_t nrec, size_t *bytes)
{
    ABK_REQUIRE(bytes && nrec >= 0, "abk_pack9_scratch_bytes: bad arguments");
    const int64_t nb = p9_blocks(nrec) > 0 ? p9_blocks(nrec) : 1;
    *bytes = abk_align_up((size_t)nb * sizeof(uint32_t), 256) + abk_scan_tmp_bytes(nb);
    return ABK_OK;
}

extern "C" int abk_pack9_count(abk_ctx *ctx, const uint8_t *data, int64_t nrec, void *scratch, size_t scratch_bytes,
                               int64_t *nheaders_h)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && nheaders_h && nrec >= 0, "abk_pack9_count: bad arguments");
    *nheaders_h = 0;
    if (nrec == 0) return ABK_OK;
    ABK_REQUIRE(data && scratch, "abk_pack9_count: null data/scratch");
    ABK_REQUIRE(((uintptr_t)data & 3) == 0 && ((uintptr_t)scratch & 255) == 0, "abk_pack9_count: data must be 4-byte and scratch 256-byte aligned");
    size_t need = 0;
    abk_pack9_scratch_bytes(nrec, &need);
    ABK_REQUIRE(scratch_bytes >= need, "abk_pack9_count: scratch too small (%zu < %zu)", scratch_bytes, need);
    const int64_t nb = p9_blocks(nrec);
    ABK_REQUIRE(nb < ((int64_t)1 << 31), "abk_pack9_count: too many records");
    uint32_t *blk = (uint32_t *)scratch;
    void *tmp = (char *)scratch + abk_align_up((size_t)nb * sizeof(uint32_t), 256);
    ABK_LAUNCH(ctx, ABK_K_MISC, pack9_count_kernel<<<(unsigned)nb, P9_BLOCK, 0, ctx->stream>>>(data, nrec, blk));
    int rc = abk_inclusive_scan_u32(ctx, blk, nb, tmp);
    if (rc) return rc;
    uint32_t total = 0;
    ABK_CHECK_CUDA(cudaMemcpyAsync(&total, blk + (nb - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    ABK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *nheaders_h = (int64_t)total;
    return ABK_OK;
}

extern "C" int abk_pack9_decode(abk_ctx *ctx, const uint8_t *data, int64_t nrec, double boxsize, double velzspace_to_kms,
                                const void *scratch, void *hdr_tab, int64_t nheaders, void *posout, void *velout, int out_f64)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && nrec >= 0 && nheaders >= 0 && nheaders <= nrec, "abk_pack9_decode: bad arguments");
    if (nrec == 0) return ABK_OK;
    ABK_REQUIRE(data && scratch && (nheaders == 0 || hdr_tab), "abk_pack9_decode: null data/scratch/header table");
    ABK_REQUIRE(((uintptr_t)data & 3) == 0, "abk_pack9_decode: data must be 4-byte aligned");
    const int64_t nb = p9_blocks(nrec);
    const uint32_t *incl = (const uint32_t *)scratch;
    if (out_f64) {
        typedef abk_pack9_cell<double> H;
        if (nheaders)
            ABK_LAUNCH(ctx, ABK_K_MISC, pack9_headers_kernel<double><<<(unsigned)nb, P9_BLOCK, 0, ctx->stream>>>(
                                            data, nrec, incl, boxsize, velzspace_to_kms, (H *)hdr_tab));
        if (posout || velout)
            ABK_LAUNCH(ctx, ABK_K_MISC, pack9_decode_kernel<double><<<(unsigned)nb, P9_BLOCK, 0, ctx->stream>>>(
                                            data, nrec, incl, (const H *)hdr_tab, (double *)posout, (double *)velout));
    } else {
        typedef abk_pack9_cell<float> H;
        if (nheaders)
            ABK_LAUNCH(ctx, ABK_K_MISC, pack9_headers_kernel<float><<<(unsigned)nb, P9_BLOCK, 0, ctx->stream>>>(
                                            data, nrec, incl, (float)boxsize, (float)velzspace_to_kms, (H *)hdr_tab));
        if (posout || velout)
            ABK_LAUNCH(ctx, ABK_K_MISC, pack9_decode_kernel<float><<<(unsigned)nb, P9_BLOCK, 0, ctx->stream>>>(
                                            data, nrec, incl, (const H *)hdr_tab, (float *)posout, (float *)velout));
    }
    return ABK_OK;
}

extern "C" int abk_unpack_pids(abk_ctx *ctx, const uint64_t *packed, int64_t N, double box, int64_t ppd, int64_t *pid,
                               void *lagr_pos, int16_t *lagr_idx, uint8_t *tagged, void *density, int out_f64)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && N >= 0 && (N == 0 || packed) && ppd > 0, "abk_unpack_pids: bad arguments");
    if (N == 0 || (!pid && !lagr_pos && !lagr_idx && !tagged && !density)) return ABK_OK;
    int64_t blocks = (N + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    if (out_f64)
        ABK_LAUNCH(ctx, ABK_K_MISC, pids_kernel<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                                        packed, N, box / (double)ppd, box / 2, pid, (double *)lagr_pos, lagr_idx, tagged, (double *)density));
    else
        ABK_LAUNCH(ctx, ABK_K_MISC, pids_kernel<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                                        packed, N, (float)(box / (double)ppd), (float)(box / 2), pid, (float *)lagr_pos, lagr_idx, tagged, (float *)density));
    return ABK_OK;
}
