// assemble.cu — multi-GPU only: on the gathering GPU, scatter `part_world` packed band buffers (rt_render_part's
// layout: the part's row bands back to back, [local_band][band_h][nx][3]) into the row-major RGB8 frame.  Pure data
// movement, HBM-bound: 2 x frame bytes (2 x 88.5 MB at 8K = ~30 us); whole 16-byte words when the row pitch allows.
#include <cstdint>

#include <cuda_runtime.h>

namespace rtb {

namespace {

template <typename W>
__global__ void assemble_bands_kernel(const unsigned char *parts, long long part_stride, int part_world, long long row_words, int ny,
                                      int band_h, unsigned char *frame) {
    // one thread per W-sized word of the frame
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= row_words * ny) return;
    const int y = (int) (idx / row_words);
    const long long xw = idx % row_words;
    const int band = y / band_h, part = band % part_world, local_band = band / part_world;
    const W *src = (const W *) (parts + (long long) part * part_stride) + ((long long) local_band * band_h + (y % band_h)) * row_words + xw;
    ((W *) frame)[idx] = *src;
}

}  // namespace

int launch_assemble(const unsigned char *parts, long long part_stride, int part_world, int nx, int ny, int band_h,
                    unsigned char *frame, cudaStream_t stream) {
    const long long row_bytes = (long long) nx * 3;
    const int threads = 256;
    const bool wide = row_bytes % 16 == 0 && part_stride % 16 == 0 && ((uintptr_t) parts % 16) == 0 && ((uintptr_t) frame % 16) == 0;
    if (wide) {
        const long long total = row_bytes / 16 * ny;
        assemble_bands_kernel<uint4><<<(unsigned) ((total + threads - 1) / threads), threads, 0, stream>>>(parts, part_stride, part_world, row_bytes / 16,
                                                                                                      ny, band_h, frame);
    } else {
        const long long total = row_bytes * ny;
        assemble_bands_kernel<unsigned char><<<(unsigned) ((total + threads - 1) / threads), threads, 0, stream>>>(parts, part_stride, part_world,
                                                                                                              row_bytes, ny, band_h, frame);
    }
    return (int) cudaGetLastError();
}

}  // namespace rtb
