// Fused filtered leaky ReLU (sm_100a): bias -> zero-insert up-sample + pad + FIR (fu) -> gain * lrelu / clamp (+ 2-bit sign codes written or read)
// -> FIR (fd) + decimate, ONE kernel, the up-sampled intermediate (up^2 x the input) living only in shared memory.
//
// Replaces filtered_lrelu_plugin.filtered_lrelu (reference src/torch_utils/ops/filtered_lrelu.cpp:16-209, kernels filtered_lrelu.cu:139-1099) for
// SEPARABLE filters (fu, fd rank 1: the configurations StyleGAN3-style layers use and BASELINE configs[3] names); rank-2 filters answer
// "unsupported" and the caller takes the generic route (upfirdn2d -> filtered_lrelu_act_ -> upfirdn2d), exactly as the reference does for
// configurations outside its own kernel table (filtered_lrelu.cpp:50-55, filtered_lrelu.py:223-229).
//
// Index contract = upfirdn2d's (upfirdn2d.cu header), applied per axis:
//   up   : c[u] = up * sum_i x[i] * fu'[i*up + p0 - u]        i in [ceil((u - p0)/up), floor((u - p0 + Fu-1)/up)] /\ [0, in)
//   act  : c[u] = sign-coded lrelu(c[u])                      (write: code 0 positive, 1 negative, 2 clamped; read: x gain, x gain*slope, x 0)
//   down : y[o] = sum_k c[o*down + k] * fd'[k]                k in [0, Fd)
// with f'[k] = f[F-1-k] (convolution) or f[k] (flip_filter).  The sign tensor is [N, C, sH, sW/4] uint8, four 2-bit codes per byte, element
// (u, v) of the intermediate at sign coordinate (u + sx, v + sy) (filtered_lrelu.cu:1136-1145; same layout as gp3d_filtered_lrelu_act).
//
// One CTA = one 32 x 32 output tile of one (n, c) plane; 256 threads; five shared-memory stages
//   sIn [INH][INW] -> up-x sB [INH][TIW] -> up-y + activation sC [TIH][TIW] -> down-x sD [TIH][32] -> down-y -> y.
// Algorithmic bytes = (numel_in + numel_out) * sizeof(T) (+ the sign bytes when written / read); the work per output is ~ (Fu/up + Fd) * 2.2 FMAs
// out of shared memory, so the kernel is shared-memory / FMA bound at 12-tap filters, not HBM bound.
#include "common.cuh"

namespace {

constexpr int TO = 32;                 // output tile edge
constexpr int kMaxF = 32;              // filter taps per axis

struct FlrParams {
    const void* x; const void* b; void* y; uint8_t* s;
    const float* fu; const float* fd;
    int N, C, xH, xW, yH, yW;
    int Fu, Fd, up, down, px0, py0;
    int sH, sW4, sx, sy, swLimit;
    float gain, slope, clamp;
    int flip, write_signs, read_signs;
    int tilesX, tilesY;
    int INW, INH, TIW, TIH;            // staged input / intermediate extents
};

__device__ __forceinline__ int fdiv(int a, int b) { int q = a / b; return (a - q * b < 0) ? q - 1 : q; }
__device__ __forceinline__ int cdiv(int a, int b) { return -fdiv(-a, b); }

template <class T>
__global__ void __launch_bounds__(256) filtered_lrelu_kernel(FlrParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float sfu[kMaxF], sfd[kMaxF];
    float* sIn = sm;
    float* sB = sIn + p.INH * p.INW;
    float* sC = sB + p.INH * p.TIW;
    float* sD = sC + p.TIH * p.TIW;
    const int tid = threadIdx.x;
    if (tid < p.Fu) sfu[tid] = p.fu[p.flip ? tid : p.Fu - 1 - tid];
    if (tid < p.Fd) sfd[tid] = p.fd[p.flip ? tid : p.Fd - 1 - tid];

    int64_t tile = blockIdx.x;
    const int tx = (int)(tile % p.tilesX); tile /= p.tilesX;
    const int ty = (int)(tile % p.tilesY); tile /= p.tilesY;
    const int nc = (int)tile;
    const int c = nc % p.C;
    const int ox0 = tx * TO, oy0 = ty * TO;
    const int u0 = ox0 * p.down, v0 = oy0 * p.down;                 // intermediate origin of this tile
    const int ix0 = cdiv(u0 - p.px0, p.up), iy0 = cdiv(v0 - p.py0, p.up);   // first input column / row that can contribute
    const T* xp = reinterpret_cast<const T*>(p.x) + (int64_t)nc * p.xH * p.xW;
    const float bias = p.b ? io_traits<T>::ld(reinterpret_cast<const T*>(p.b) + c) : 0.f;

    // A: input tile (+ bias); outside the image: 0 (the padding region of the zero-inserted signal carries no bias)
    for (int e = tid; e < p.INH * p.INW; e += 256) {
        const int r = e / p.INW, q = e - r * p.INW;
        const int iy = iy0 + r, ix = ix0 + q;
        sIn[e] = (iy >= 0 && iy < p.xH && ix >= 0 && ix < p.xW) ? io_traits<T>::ld(xp + (int64_t)iy * p.xW + ix) + bias : 0.f;
    }
    __syncthreads();
    // B: up-sample along x
    for (int e = tid; e < p.INH * p.TIW; e += 256) {
        const int r = e / p.TIW, q = e - r * p.TIW;
        const int u = u0 + q;
        int ilo = cdiv(u - p.px0, p.up), ihi = fdiv(u - p.px0 + p.Fu - 1, p.up);
        const float* row = sIn + r * p.INW - ix0;
        float acc = 0.f;
        for (int i = ilo; i <= ihi; i++) acc = fmaf(row[i], sfu[i * p.up + p.px0 - u], acc);      // staged zeros cover i outside the image
        sB[e] = acc;
    }
    __syncthreads();
    // C: up-sample along y, gain up^2, activation with sign codes; a thread owns 4 consecutive u (one sign byte)
    const float upgain = (float)(p.up * p.up);
    const int q4n = (p.TIW + 3) >> 2;
    uint8_t* sp = p.s ? p.s + (int64_t)nc * p.sH * p.sW4 : nullptr;
    const int own_u_end = (tx == p.tilesX - 1) ? 0x7fffffff : u0 + TO * p.down;      // sign bytes are written by the tile that owns the column range
    const int own_v_end = (ty == p.tilesY - 1) ? 0x7fffffff : v0 + TO * p.down;
    for (int e = tid; e < p.TIH * q4n; e += 256) {
        const int r = e / q4n, q = (e - r * q4n) * 4;
        const int v = v0 + r;
        const int jlo = cdiv(v - p.py0, p.up), jhi = fdiv(v - p.py0 + p.Fu - 1, p.up);
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = jlo; j <= jhi; j++) {
            const float w = sfu[j * p.up + p.py0 - v];
            const float* src = sB + (j - iy0) * p.TIW + q;
#pragma unroll
            for (int k = 0; k < 4; k++) if (q + k < p.TIW) a[k] = fmaf(src[k], w, a[k]);
        }
        const int sv = v + p.sy;
        const bool srow = sp && sv >= 0 && sv < p.sH;
        uint32_t wcode = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (q + k >= p.TIW) break;
            float val = a[k] * upgain;
            if (p.read_signs) {
                const int su = u0 + q + k + p.sx;
                uint32_t code = 0;
                if (srow && su >= 0 && (su >> 2) < p.sW4) code = (sp[(int64_t)sv * p.sW4 + (su >> 2)] >> ((su & 3) * 2)) & 3u;
                val *= (code == 0) ? p.gain : (code == 1) ? p.gain * p.slope : 0.f;
            } else {
                uint32_t code = 0;
                if (val < 0.f) { val *= p.slope; code = 1; }
                val *= p.gain;
                if (p.clamp >= 0.f && fabsf(val) > p.clamp) { val = copysignf(p.clamp, val); code = 2; }
                wcode |= code << (2 * k);
            }
            sC[r * p.TIW + q + k] = val;
        }
        if (p.write_signs && srow && v < own_v_end) {
            const int su = u0 + q + p.sx;                          // multiple of 4 (launcher checks sx % 4 == 0; u0, q are multiples of 4)
            if (u0 + q < own_u_end && su >= 0 && (su >> 2) < p.swLimit) sp[(int64_t)sv * p.sW4 + (su >> 2)] = (uint8_t)wcode;
        }
    }
    __syncthreads();
    // D: down-sample along x
    for (int e = tid; e < p.TIH * TO; e += 256) {
        const int r = e / TO, q = e - r * TO;
        const float* src = sC + r * p.TIW + q * p.down;
        float acc = 0.f;
        for (int k = 0; k < p.Fd; k++) acc = fmaf(src[k], sfd[k], acc);
        sD[e] = acc;
    }
    __syncthreads();
    // E: down-sample along y, store
    T* yp = reinterpret_cast<T*>(p.y) + (int64_t)nc * p.yH * p.yW;
    for (int e = tid; e < TO * TO; e += 256) {
        const int r = e / TO, q = e - r * TO;
        const int oy = oy0 + r, ox = ox0 + q;
        if (oy >= p.yH || ox >= p.yW) continue;
        const float* src = sD + (r * p.down) * TO + q;
        float acc = 0.f;
        for (int k = 0; k < p.Fd; k++) acc = fmaf(src[k * TO], sfd[k], acc);
        io_traits<T>::st(yp + (int64_t)oy * p.yW + ox, acc);
    }
}


// ------------------------------------------------------------------------------------------------------------------------------------------------
// Register-blocked variant for the filter families StyleGAN3-style layers use (taps = 6 x rate, or 1): compile-time rates / lengths, every stage a
// SLIDING WINDOW held in registers -- a work item owns one polyphase lane (fixed coefficient set, loaded once) and walks along the filter axis, so an
// output costs FU/UP (or FD) FMAs for ONE (or DOWN) shared-memory loads instead of two loads per FMA.  Row pitches are odd (bank-conflict-free
// column walks); the down-x buffer aliases the dead input / up-x buffers.  Same arithmetic contract as the generic kernel above.
template <class T, int UP, int DOWN, int FU, int FD>
__global__ void __launch_bounds__(256) filtered_lrelu_fast_kernel(FlrParams p) {
    constexpr int NT = FU / UP;                                     // taps per up-sampling phase (FU % UP == 0)
    constexpr int TI = (TO - 1) * DOWN + FD;                        // intermediate rows / columns a tile needs
    constexpr int TIP = ((TI + 3) & ~3) | 1;                        // pitch: whole sign bytes, odd
    constexpr int IN = (((TI + 3) & ~3) + FU - 1) / UP + 2;         // staged input rows / columns
    constexpr int INP = IN | 1;
    constexpr int DP = TO + 1;                                      // pitch of the down-x buffer
    extern __shared__ __align__(16) float sm[];
    __shared__ float sfu[FU], sfd[FD];
    float* sIn = sm;                        // [IN][INP]
    float* sB = sIn + IN * INP;             // [IN][TIP]
    float* sC = sB + IN * TIP;              // [TI][TIP]
    float* sD = sm;                         // [TI][DP]   (aliases sIn / sB, dead by then)
    static_assert(TI * DP <= IN * INP + IN * TIP, "down-x buffer must fit into the dead input + up-x buffers");
    const int tid = threadIdx.x;
    if (tid < FU) sfu[tid] = p.fu[p.flip ? tid : FU - 1 - tid];
    if (tid < FD) sfd[tid] = p.fd[p.flip ? tid : FD - 1 - tid];

    int64_t tile = blockIdx.x;
    const int tx = (int)(tile % p.tilesX); tile /= p.tilesX;
    const int ty = (int)(tile % p.tilesY); tile /= p.tilesY;
    const int nc = (int)tile;
    const int c = nc % p.C;
    const int ox0 = tx * TO, oy0 = ty * TO;
    const int u0 = ox0 * DOWN, v0 = oy0 * DOWN;
    const int ix0 = cdiv(u0 - p.px0, UP), iy0 = cdiv(v0 - p.py0, UP);
    const T* xp = reinterpret_cast<const T*>(p.x) + (int64_t)nc * p.xH * p.xW;
    const float bias = p.b ? io_traits<T>::ld(reinterpret_cast<const T*>(p.b) + c) : 0.f;

    // A: input tile (+ bias), zero outside the image
    for (int e = tid; e < IN * IN; e += 256) {
        const int r = e / IN, q = e - r * IN;
        const int iy = iy0 + r, ix = ix0 + q;
        sIn[r * INP + q] = (iy >= 0 && iy < p.xH && ix >= 0 && ix < p.xW) ? io_traits<T>::ld(xp + (int64_t)iy * p.xW + ix) + bias : 0.f;
    }
    __syncthreads();

    // B: up-sample along x.  Work item = (input row r, phase lane rho, column segment): columns q = rho + m * UP share one coefficient set and slide
    // one input per output.
    {
        constexpr int NCOL = (TIP - 1 + UP - 1) / UP;               // outputs per (row, phase)
        constexpr int BSEG = 4;
        constexpr int MPER = (NCOL + BSEG - 1) / BSEG;
        for (int wi = tid; wi < IN * UP * BSEG; wi += 256) {
            const int r = wi % IN, rs = wi / IN;                    // lanes <-> consecutive rows (odd pitches: conflict-free)
            const int rho = rs % UP, m0 = (rs / UP) * MPER;
            float cf[NT], w[NT];
            const int ilo0 = cdiv(u0 + rho - p.px0, UP);
            const int k0 = ilo0 * UP + p.px0 - (u0 + rho);         // tap of the first contributing sample, in [0, UP)
#pragma unroll
            for (int t = 0; t < NT; t++) cf[t] = sfu[k0 + t * UP];
            const float* src = sIn + r * INP + (ilo0 - ix0) + m0;
#pragma unroll
            for (int t = 0; t < NT - 1; t++) w[t + 1] = src[t];
            float* dst = sB + r * TIP + rho + m0 * UP;
#pragma unroll 4
            for (int m = 0; m < MPER; m++) {
                if (rho + (m0 + m) * UP >= TIP - 1) break;
#pragma unroll
                for (int t = 0; t < NT - 1; t++) w[t] = w[t + 1];
                w[NT - 1] = src[m + NT - 1];
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int t = 0; t < NT; t += 2) { a0 = fmaf(w[t], cf[t], a0); if (t + 1 < NT) a1 = fmaf(w[t + 1], cf[t + 1], a1); }
                dst[m * UP] = a0 + a1;
            }
        }
    }
    __syncthreads();

    // C: up-sample along y + activation.  Work item = (column q, phase lane rho, row segment); lanes <-> consecutive columns, so the four lanes of a
    // sign byte combine their 2-bit codes with shuffles.
    {
        constexpr int QW = (TIP - 1 + 31) / 32 * 32;                // columns rounded to whole warps
        constexpr int SEG = 4;                                      // row segments per (column, phase)
        constexpr int ROWS_PER = (TI + UP * SEG - 1) / (UP * SEG);  // outputs per work item
        const float upgain = (float)(UP * UP);
        uint8_t* sp = p.s ? p.s + (int64_t)nc * p.sH * p.sW4 : nullptr;
        const int own_u_end = (tx == p.tilesX - 1) ? 0x7fffffff : u0 + TO * DOWN;
        const int own_v_end = (ty == p.tilesY - 1) ? 0x7fffffff : v0 + TO * DOWN;
        for (int wi = tid; wi < QW * UP * SEG; wi += 256) {         // QW is a multiple of 32: whole warps stay together
            const int q = wi % QW, rs = wi / QW;
            const int rho = rs % UP, seg = rs / UP;
            const bool colok = q < TIP - 1;
            const int m0 = seg * ROWS_PER;                          // first output index of this lane: row = rho + (m0 + m) * UP
            float cf[NT], w[NT];
            const int jlo0 = cdiv(v0 + rho - p.py0, UP);
            const int k0 = jlo0 * UP + p.py0 - (v0 + rho);
#pragma unroll
            for (int t = 0; t < NT; t++) cf[t] = sfu[k0 + t * UP] * upgain;
            const float* src = sB + (jlo0 - iy0 + m0) * TIP + (colok ? q : 0);
#pragma unroll
            for (int t = 0; t < NT - 1; t++) w[t + 1] = src[t * TIP];
            for (int m = 0; m < ROWS_PER; m++) {
                const int r = rho + (m0 + m) * UP;
                const bool rowok = r < TI;                          // warp-uniform (rho, seg, m are)
                if (!rowok) break;
#pragma unroll
                for (int t = 0; t < NT - 1; t++) w[t] = w[t + 1];
                w[NT - 1] = src[(m + NT - 1) * TIP];
                float val = 0.f, val1 = 0.f;
#pragma unroll
                for (int t = 0; t < NT; t += 2) { val = fmaf(w[t], cf[t], val); if (t + 1 < NT) val1 = fmaf(w[t + 1], cf[t + 1], val1); }
                val += val1;
                const int v = v0 + r, sv = v + p.sy;
                const bool srow = sp && sv >= 0 && sv < p.sH;
                uint32_t code = 0;
                if (p.read_signs) {
                    const int su = u0 + q + p.sx;
                    if (srow && su >= 0 && (su >> 2) < p.sW4) code = (sp[(int64_t)sv * p.sW4 + (su >> 2)] >> ((su & 3) * 2)) & 3u;
                    val *= (code == 0) ? p.gain : (code == 1) ? p.gain * p.slope : 0.f;
                } else {
                    if (val < 0.f) { val *= p.slope; code = 1; }
                    val *= p.gain;
                    if (p.clamp >= 0.f && fabsf(val) > p.clamp) { val = copysignf(p.clamp, val); code = 2; }
                }
                if (colok) sC[r * TIP + q] = val;
                if (p.write_signs) {                                // warp-uniform branch
                    if (!colok) code = 0;
                    uint32_t byte = code;
                    byte |= __shfl_down_sync(0xffffffffu, code, 1) << 2;
                    byte |= __shfl_down_sync(0xffffffffu, code, 2) << 4;
                    byte |= __shfl_down_sync(0xffffffffu, code, 3) << 6;
                    const int su = u0 + q + p.sx;                   // multiple of 4 for the byte owner (q % 4 == 0)
                    if ((q & 3) == 0 && srow && v < own_v_end && u0 + q < own_u_end && su >= 0 && (su >> 2) < p.swLimit)
                        sp[(int64_t)sv * p.sW4 + (su >> 2)] = (uint8_t)byte;
                }
            }
        }
    }
    __syncthreads();

    // D: down-sample along x.  Work item = (row r, output segment of 4): window of FD values sliding by DOWN.
    constexpr int OS = 4;
    for (int wi = tid; wi < TI * (TO / OS); wi += 256) {
        const int r = wi % TI, seg = wi / TI;                       // lanes <-> consecutive rows (odd pitch)
        float cf[FD], w[FD];
#pragma unroll
        for (int k = 0; k < FD; k++) cf[k] = sfd[k];
        const float* src = sC + r * TIP + seg * OS * DOWN;
#pragma unroll
        for (int k = 0; k < FD - DOWN; k++) w[k + DOWN] = src[k];
#pragma unroll
        for (int o = 0; o < OS; o++) {
#pragma unroll
            for (int k = 0; k < FD - DOWN; k++) w[k] = w[k + DOWN];
#pragma unroll
            for (int k = 0; k < DOWN; k++) w[FD - DOWN + k] = src[o * DOWN + FD - DOWN + k];
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < FD; k += 2) { a0 = fmaf(w[k], cf[k], a0); if (k + 1 < FD) a1 = fmaf(w[k + 1], cf[k + 1], a1); }
            sD[r * DP + seg * OS + o] = a0 + a1;
        }
    }
    __syncthreads();

    // E: down-sample along y, store.  Work item = (column q, row segment of 4).
    T* yp = reinterpret_cast<T*>(p.y) + (int64_t)nc * p.yH * p.yW;
    for (int wi = tid; wi < TO * (TO / OS); wi += 256) {
        const int q = wi % TO, seg = wi / TO;
        float cf[FD], w[FD];
#pragma unroll
        for (int k = 0; k < FD; k++) cf[k] = sfd[k];
        const float* src = sD + (seg * OS * DOWN) * DP + q;
#pragma unroll
        for (int k = 0; k < FD - DOWN; k++) w[k + DOWN] = src[k * DP];
        const int ox = ox0 + q;
#pragma unroll
        for (int o = 0; o < OS; o++) {
#pragma unroll
            for (int k = 0; k < FD - DOWN; k++) w[k] = w[k + DOWN];
#pragma unroll
            for (int k = 0; k < DOWN; k++) w[FD - DOWN + k] = src[(o * DOWN + FD - DOWN + k) * DP];
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < FD; k += 2) { a0 = fmaf(w[k], cf[k], a0); if (k + 1 < FD) a1 = fmaf(w[k + 1], cf[k + 1], a1); }
            const int oy = oy0 + seg * OS + o;
            if (oy < p.yH && ox < p.yW) io_traits<T>::st(yp + (int64_t)oy * p.yW + ox, a0 + a1);
        }
    }
}

template <class T, int UP, int DOWN, int FU, int FD>
int launch_fast(const FlrParams& p, int64_t tiles, cudaStream_t st) {
    constexpr int TI = (TO - 1) * DOWN + FD;
    constexpr int TIP = ((TI + 3) & ~3) | 1;
    constexpr int IN = (((TI + 3) & ~3) + FU - 1) / UP + 2;
    constexpr int INP = IN | 1;
    const size_t smem = (size_t)(IN * INP + IN * TIP + TI * TIP) * sizeof(float);
    auto kern = filtered_lrelu_fast_kernel<T, UP, DOWN, FU, FD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("filtered_lrelu: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    kern<<<(unsigned)tiles, 256, smem, st>>>(p);
    return -1000;     // launched (the caller checks cudaGetLastError)
}

// Kernel table: (up, down, fu taps, fd taps).  StyleGAN3's layers use taps = 6 x rate (or a single tap at rate 1).
template <class T>
int dispatch_fast(const FlrParams& p, int64_t tiles, cudaStream_t st) {
#define GP3D_FLR_CASE(U, D, A, B) if (p.up == U && p.down == D && p.Fu == A && p.Fd == B) return launch_fast<T, U, D, A, B>(p, tiles, st);
    GP3D_FLR_CASE(2, 2, 12, 12)
    GP3D_FLR_CASE(4, 2, 24, 12)
    GP3D_FLR_CASE(2, 4, 12, 24)
    GP3D_FLR_CASE(2, 1, 12, 1)
    GP3D_FLR_CASE(1, 2, 1, 12)
    GP3D_FLR_CASE(1, 1, 6, 6)
#undef GP3D_FLR_CASE
    return 0;         // not in the table
}

}  // namespace

extern "C" int gp3d_filtered_lrelu(const void* x, const float* fu, const float* fd, const void* b, uint8_t* s, void* y, int dtype,
                                   int N, int C, int xH, int xW, int yH, int yW, int fu_taps, int fd_taps, int up, int down, int px0, int py0,
                                   int sH, int sW4, int sx, int sy, int sw_limit, float gain, float slope, float clamp, int flip_filter,
                                   int write_signs, int read_signs, void* stream) {
    GP3D_CHECK_ARG(x && fu && fd && y && N >= 1 && C >= 1 && xH >= 1 && xW >= 1 && yH >= 1 && yW >= 1, "filtered_lrelu: bad arguments");
    GP3D_CHECK_ARG(up >= 1 && down >= 1 && fu_taps >= 1 && fd_taps >= 1, "filtered_lrelu: up, down and the filter lengths must be at least 1");
    GP3D_CHECK_ARG(!(write_signs && read_signs), "filtered_lrelu: signs are either written or read");
    GP3D_CHECK_ARG(!(write_signs || read_signs) || (s && sH >= 1 && sW4 >= 1), "filtered_lrelu: bad sign tensor geometry");
    GP3D_CHECK_ARG(dtype == GP3D_F32 || dtype == GP3D_F16, "filtered_lrelu: x must be float32 or float16");
    if (fu_taps > kMaxF || fd_taps > kMaxF || fu_taps < up || fd_taps < down || (write_signs && (sx & 3) != 0)) {
        gp3d_set_error("filtered_lrelu: configuration outside the fused kernel (taps <= %d, taps >= rate, sign x-offset %% 4 == 0 when writing)", kMaxF);
        return GP3D_E_UNSUPPORTED;
    }
    FlrParams p{};
    p.x = x; p.b = b; p.y = y; p.s = (write_signs || read_signs) ? s : nullptr; p.fu = fu; p.fd = fd;
    p.N = N; p.C = C; p.xH = xH; p.xW = xW; p.yH = yH; p.yW = yW;
    p.Fu = fu_taps; p.Fd = fd_taps; p.up = up; p.down = down; p.px0 = px0; p.py0 = py0;
    p.sH = sH; p.sW4 = sW4; p.sx = sx; p.sy = sy; p.swLimit = sw_limit;
    p.gain = gain; p.slope = slope; p.clamp = (clamp >= 0.f && clamp < 3.0e38f) ? clamp : -1.f;
    p.flip = flip_filter; p.write_signs = write_signs; p.read_signs = read_signs;
    p.tilesX = (yW + TO - 1) / TO; p.tilesY = (yH + TO - 1) / TO;
    p.TIW = (TO - 1) * down + fd_taps; p.TIH = p.TIW;
    p.TIW = (p.TIW + 3) & ~3;                                       // whole sign bytes per row
    p.INW = (p.TIW + fu_taps - 1) / up + 2; p.INH = (p.TIH + fu_taps - 1) / up + 2;
    const size_t smem = (size_t)(p.INH * p.INW + p.INH * p.TIW + p.TIH * p.TIW + p.TIH * TO) * sizeof(float);
    if (smem > 200 * 1024) { gp3d_set_error("filtered_lrelu: tile needs %zu B of shared memory", smem); return GP3D_E_UNSUPPORTED; }
    const int64_t tiles = (int64_t)p.tilesX * p.tilesY * N * C;
    GP3D_CHECK_ARG(tiles < 2147483647LL, "filtered_lrelu: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int rc = (dtype == GP3D_F32) ? dispatch_fast<float>(p, tiles, st) : dispatch_fast<__half>(p, tiles, st);
        if (rc == -1000) GP3D_RETURN_LAUNCH();
        if (rc != 0) return rc;
    }
    auto kern = (dtype == GP3D_F32) ? filtered_lrelu_kernel<float> : filtered_lrelu_kernel<__half>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("filtered_lrelu: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    kern<<<(unsigned)tiles, 256, smem, st>>>(p);
    GP3D_RETURN_LAUNCH();
}
