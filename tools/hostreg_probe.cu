// hostreg_probe.cu -- can the file pipeline DMA straight into / out of file mappings?  Times cudaHostRegister on a tmpfs
// output mapping (MAP_SHARED, fresh and populated pages) and on a read-only input mapping, in pieces of 64 MiB and as
// a whole, and the copies that follow.   nvcc -O2 -o /tmp/hostreg_probe tools/hostreg_probe.cu && /tmp/hostreg_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("%s -> %s\n", #x, cudaGetErrorString(e)); } } while (0)

int main() {
    const size_t n = size_t(2) << 30, piece = size_t(64) << 20;
    CK(cudaFree(0));
    void *d = nullptr;
    CK(cudaMalloc(&d, n));
    const char *path = "/dev/shm/blt_hostreg_probe.bin";
    int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0 || ftruncate(fd, off_t(n)) != 0) { std::printf("cannot create %s\n", path); return 1; }
    unsigned char *m = static_cast<unsigned char *>(mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
    double t0 = now();
    CK(cudaHostRegister(m, n / 2, cudaHostRegisterDefault));   // fresh (unallocated) pages
    double t1 = now();
    std::printf("register 1 GiB of FRESH shared tmpfs mapping: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    t0 = now();
    CK(cudaMemcpy(m, d, n / 2, cudaMemcpyDeviceToHost));
    t1 = now();
    std::printf("  D2H into it: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    t0 = now();
    CK(cudaHostUnregister(m));
    t1 = now();
    std::printf("  unregister: %.3f s\n", t1 - t0);
    // second half: fallocate + populate first, then register piecewise
    t0 = now();
    if (fallocate(fd, 0, off_t(n / 2), off_t(n / 2)) != 0) std::printf("fallocate failed\n");
    madvise(m + n / 2, n / 2, 23 /* MADV_POPULATE_WRITE */);
    t1 = now();
    std::printf("fallocate + populate 1 GiB: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    t0 = now();
    for (size_t off = n / 2; off < n; off += piece) CK(cudaHostRegister(m + off, piece, cudaHostRegisterDefault));
    t1 = now();
    std::printf("register 1 GiB of POPULATED mapping in 64 MiB pieces: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    t0 = now();
    CK(cudaMemcpy(m + n / 2, d, n / 2, cudaMemcpyDeviceToHost));
    t1 = now();
    std::printf("  D2H into it: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    t0 = now();
    for (size_t off = n / 2; off < n; off += piece) CK(cudaHostUnregister(m + off));
    t1 = now();
    std::printf("  unregister pieces: %.3f s\n", t1 - t0);
    // input side: read-only private mapping of the same file
    unsigned char *r = static_cast<unsigned char *>(mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0));
    t0 = now();
    cudaError_t e = cudaHostRegister(r, n / 2, cudaHostRegisterReadOnly);
    t1 = now();
    std::printf("register 1 GiB of READ-ONLY private mapping: %s, %.3f s (%.2f GB/s)\n", cudaGetErrorString(e), t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    if (e == cudaSuccess) {
        t0 = now();
        CK(cudaMemcpy(d, r, n / 2, cudaMemcpyHostToDevice));
        t1 = now();
        std::printf("  H2D from it: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
        CK(cudaHostUnregister(r));
    } else {
        cudaGetLastError();
    }
    // baseline: pageable copies straight from / into the mappings (the driver's own staging)
    t0 = now();
    CK(cudaMemcpy(d, r + n / 2, n / 2, cudaMemcpyHostToDevice));
    t1 = now();
    std::printf("pageable H2D from the mapping: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    t0 = now();
    CK(cudaMemcpy(m, d, n / 2, cudaMemcpyDeviceToHost));
    t1 = now();
    std::printf("pageable D2H into the mapping: %.3f s (%.2f GB/s)\n", t1 - t0, double(n / 2) / (t1 - t0) / 1e9);
    munmap(r, n); munmap(m, n); close(fd); unlink(path);
    return 0;
}
