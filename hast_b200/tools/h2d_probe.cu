// h2d_probe.cu -- what can this BOX move host -> device when 1, 2, 4, 8 GPUs copy at once?
//
// The e2e leg of bench.py (hast_submit_batch: pinned host buffers, cudaMemcpyAsync) runs at the PCIe rate of one
// GPU when it is alone and at about half of that per GPU when eight copy together.  This probe measures that
// ceiling with nothing but cudaMemcpyAsync, and tries the placements that could raise it:
//   default   cudaHostAlloc(portable) from the launching thread            (what hast_host_alloc does)
//   local     a thread pinned to the CPUs of the GPU's NUMA node allocates, first-touches and cudaHostRegister()s
//             its buffer, and issues the copies from there
//   wc        cudaHostAllocWriteCombined
// Output: one JSON object on stdout, {"per_gpu_gbs": {"1": .., "2": .., ...}, "variants": {...}, "topology": [...]}.
// Measurement tool, not on the classification path.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static std::string read_line(const std::string& path) {
    std::ifstream f(path);
    std::string s;
    std::getline(f, s);
    return s;
}

static std::string pci_path(int dev) {
    char bus[32] = {0};
    cudaDeviceGetPCIBusId(bus, sizeof bus, dev);
    std::string b(bus);
    for (auto& c : b) c = (char)tolower(c);
    return "/sys/bus/pci/devices/" + b;
}

static bool bind_to_cpulist(const std::string& list) {     // "0-15,32-47"
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    size_t i = 0;
    while (i < list.size()) {
        char* end = nullptr;
        long a = strtol(list.c_str() + i, &end, 10), b = a;
        if (end == list.c_str() + i) break;
        i = (size_t)(end - list.c_str());
        if (i < list.size() && list[i] == '-') { b = strtol(list.c_str() + i + 1, &end, 10); i = (size_t)(end - list.c_str()); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, &set); ++n; }
        if (i < list.size() && list[i] == ',') ++i;
    }
    return n > 0 && sched_setaffinity(0, sizeof set, &set) == 0;
}

enum Variant { kDefault = 0, kLocal = 1, kWc = 2 };

// every participating GPU copies `bytes` `reps` times after a common start; returns per-GPU GB/s (wall of the slowest)
static double run(int n_gpu, Variant v, size_t bytes, double seconds, std::vector<double>* each) {
    std::vector<std::thread> th;
    std::atomic<int> ready{0}, go{0};
    std::vector<double> rate((size_t)n_gpu, 0.0);
    for (int g = 0; g < n_gpu; ++g)
        th.emplace_back([&, g] {
            cudaSetDevice(g);
            if (v == kLocal) {
                const std::string node = read_line(pci_path(g) + "/numa_node");
                if (!node.empty() && node != "-1") bind_to_cpulist(read_line("/sys/devices/system/node/node" + node + "/cpulist"));
            }
            void* h = nullptr;
            bool registered = false;
            if (v == kLocal) {
                h = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                memset(h, 1, bytes);                                  // first touch on this thread's node
                registered = cudaHostRegister(h, bytes, cudaHostRegisterPortable) == cudaSuccess;
            } else {
                cudaHostAlloc(&h, bytes, cudaHostAllocPortable | (v == kWc ? cudaHostAllocWriteCombined : 0));
                memset(h, 1, bytes);
            }
            void* d = nullptr;
            cudaMalloc(&d, bytes);
            cudaStream_t s;
            cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
            cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);  // warm-up
            cudaStreamSynchronize(s);
            ready.fetch_add(1);
            while (!go.load()) std::this_thread::yield();
            const double t0 = now();
            size_t moved = 0;
            while (now() - t0 < seconds) {
                for (int r = 0; r < 4; ++r) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
                cudaStreamSynchronize(s);
                moved += 4 * bytes;
            }
            rate[(size_t)g] = (double)moved / (now() - t0) / 1e9;
            cudaStreamDestroy(s);
            cudaFree(d);
            if (v == kLocal) { if (registered) cudaHostUnregister(h); munmap(h, bytes); } else cudaFreeHost(h);
        });
    while (ready.load() < n_gpu) std::this_thread::yield();
    go.store(1);
    for (auto& t : th) t.join();
    double mn = 1e30;
    for (double r : rate) mn = std::min(mn, r);
    if (each) *each = rate;
    return mn;
}

int main(int argc, char** argv) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const size_t bytes = (size_t)(argc > 1 ? atol(argv[1]) : 256) << 20;
    const double seconds = argc > 2 ? atof(argv[2]) : 1.0;
    printf("{\"buffer_mib\": %zu, \"seconds_per_point\": %.2f, \"gpus_visible\": %d, \"host_cpus\": %ld, \"topology\": [",
           bytes >> 20, seconds, n_dev, sysconf(_SC_NPROCESSORS_ONLN));
    for (int g = 0; g < n_dev; ++g) {
        const std::string p = pci_path(g);
        const std::string node = read_line(p + "/numa_node");
        printf("%s{\"gpu\": %d, \"pci\": \"%s\", \"numa_node\": \"%s\", \"local_cpus\": \"%s\", \"link_speed\": \"%s\", \"link_width\": \"%s\"}",
               g ? ", " : "", g, p.substr(p.rfind('/') + 1).c_str(), node.c_str(), read_line(p + "/local_cpulist").c_str(),
               read_line(p + "/current_link_speed").c_str(), read_line(p + "/current_link_width").c_str());
    }
    printf("], \"numa_nodes_online\": \"%s\", \"variants\": {", read_line("/sys/devices/system/node/online").c_str());
    const char* names[3] = {"default", "local", "wc"};
    double def_rate[9] = {0};
    for (int v = 0; v < 3; ++v) {
        printf("%s\"%s\": {", v ? ", " : "", names[v]);
        bool first = true;
        for (int n : {1, 2, 4, 8}) {
            if (n > n_dev) break;
            std::vector<double> each;
            const double r = run(n, (Variant)v, bytes, seconds, &each);
            if (v == 0) def_rate[n] = r;
            double sum = 0;
            for (double e : each) sum += e;
            printf("%s\"%d\": {\"per_gpu_min_gbs\": %.2f, \"aggregate_gbs\": %.2f}", first ? "" : ", ", n, r, sum);
            first = false;
        }
        printf("}");
    }
    printf("}, \"per_gpu_gbs\": {");
    bool first = true;
    for (int n : {1, 2, 4, 8}) {
        if (n > n_dev) break;
        printf("%s\"%d\": %.2f", first ? "" : ", ", n, def_rate[n]);
        first = false;
    }
    printf("}, \"how\": \"cudaMemcpyAsync of a pinned %zu MiB buffer, back to back for %.1f s per point, all GPUs started together; per_gpu_gbs = the slowest GPU's rate with the default placement (cudaHostAlloc portable)\"}\n",
           bytes >> 20, seconds);
    return 0;
}
