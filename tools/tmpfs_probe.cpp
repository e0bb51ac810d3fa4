// tmpfs_probe.cpp -- how fast can fresh tmpfs pages be produced?  Decides how the file pipeline writes its
// output ("file pipeline").  g++ -O2 -pthread -o blt_b200/lib/tmpfs_probe tools/tmpfs_probe.cpp
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    const size_t bytes = (argc > 1 ? size_t(atoll(argv[1])) : 4) << 30;
    const char *path = "/dev/shm/blt_probe.bin";
    std::vector<unsigned char> src(64 << 20, 0x5a);
    auto fresh = [&](bool map, unsigned char **m) {
        int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (ftruncate(fd, off_t(bytes)) != 0) perror("ftruncate");
        *m = nullptr;
        if (map) *m = static_cast<unsigned char *>(mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
        return fd;
    };
    auto par = [&](int T, auto fn) {
        std::vector<std::thread> th;
        const double t0 = now();
        for (int t = 0; t < T; ++t) th.emplace_back(fn, t, T);
        for (auto &x : th) x.join();
        return now() - t0;
    };
    const size_t piece = 16 << 20;
    for (int T : {1, 2, 4, 8, 16, 24}) {
        unsigned char *m;
        int fd = fresh(true, &m);
        double s = par(T, [&](int t, int TT) { for (size_t o = size_t(t) * piece; o < bytes; o += size_t(TT) * piece) memcpy(m + o, src.data(), piece); });
        printf("memcpy into fresh mapping, %2d threads: %.2f GB/s\n", T, bytes / s / 1e9);
        // second pass: pages exist and are mapped
        s = par(T, [&](int t, int TT) { for (size_t o = size_t(t) * piece; o < bytes; o += size_t(TT) * piece) memcpy(m + o, src.data(), piece); });
        printf("memcpy into populated mapping, %2d threads: %.2f GB/s\n", T, bytes / s / 1e9);
        munmap(m, bytes); close(fd);
    }
    for (int T : {1, 4, 8, 16}) {
        unsigned char *m;
        int fd = fresh(true, &m);
        double s = par(T, [&](int t, int TT) { for (size_t o = size_t(t) * piece; o < bytes; o += size_t(TT) * piece) if (madvise(m + o, piece, MADV_POPULATE_WRITE) != 0) { perror("madvise"); return; } });
        printf("MADV_POPULATE_WRITE, %2d threads: %.2f GB/s\n", T, bytes / s / 1e9);
        munmap(m, bytes); close(fd);
    }
    {
        unsigned char *m;
        int fd = fresh(false, &m);
        double t0 = now();
        if (fallocate(fd, 0, 0, off_t(bytes)) != 0) perror("fallocate");
        printf("fallocate, 1 thread: %.2f GB/s\n", bytes / (now() - t0) / 1e9);
        m = static_cast<unsigned char *>(mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
        double s = par(8, [&](int t, int TT) { for (size_t o = size_t(t) * piece; o < bytes; o += size_t(TT) * piece) memcpy(m + o, src.data(), piece); });
        printf("memcpy into fallocated (unmapped) pages, 8 threads: %.2f GB/s\n", bytes / s / 1e9);
        munmap(m, bytes); close(fd);
    }
    for (int T : {1, 4, 8}) {
        unsigned char *m;
        int fd = fresh(false, &m);
        double s = par(T, [&](int t, int TT) { for (size_t o = size_t(t) * piece; o < bytes; o += size_t(TT) * piece) if (pwrite(fd, src.data(), piece, off_t(o)) != ssize_t(piece)) { perror("pwrite"); return; } });
        printf("pwrite, %2d threads: %.2f GB/s\n", T, bytes / s / 1e9);
        close(fd);
    }
    {   // anonymous memory for comparison (what a Vec<u8> result costs the reference)
        unsigned char *m = static_cast<unsigned char *>(mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
        double s = par(8, [&](int t, int TT) { for (size_t o = size_t(t) * piece; o < bytes; o += size_t(TT) * piece) memcpy(m + o, src.data(), piece); });
        printf("memcpy into fresh anonymous memory, 8 threads: %.2f GB/s\n", bytes / s / 1e9);
        munmap(m, bytes);
    }
    unlink(path);
    return 0;
}
