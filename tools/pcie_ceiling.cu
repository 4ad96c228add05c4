// Bare pinned-memory copy ceiling of the platform: what cudaMemcpyAsync alone reaches between page-locked host memory
// and HBM, one direction at a time and both at once, with N processes (one per GPU) running concurrently. bench.py's
// `e2e` numbers are quoted against these figures (profiles/*pcie_ceiling*.json): they tell the platform from the engine.
//
//   nvcc -O2 -o tools/pcie_ceiling tools/pcie_ceiling.cu
//   tools/pcie_ceiling --device D [--h2d-mb 526] [--d2h-mb 3146] [--reps 5] [--start-at EPOCH_SECONDS] [--numa]
//
// Default sizes are one step of the headline workload (8192 HCA streams in, 8192 WAV images out). With --start-at all
// processes of a multi-GPU run spin until the same wall-clock second so their copies overlap. Prints one JSON line.
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static double now_s() {
    return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// pin this process to the CPUs of the GPU's NUMA node (so that the page-locked buffers are node-local)
static int bind_numa(int device, char* note, size_t note_len) {
    char bus[32] = {};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
    int node = -1;
    if (!(f >> node) || node < 0) { snprintf(note, note_len, "numa_node unknown"); return -1; }
    std::ifstream c("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!(c >> list)) return -1;
    cpu_set_t set;
    CPU_ZERO(&set);
    size_t pos = 0;
    while (pos < list.size()) {
        size_t comma = list.find(',', pos);
        std::string part = list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        int a = 0, b = 0;
        if (sscanf(part.c_str(), "%d-%d", &a, &b) == 2) { for (int k = a; k <= b; k++) CPU_SET(k, &set); }
        else if (sscanf(part.c_str(), "%d", &a) == 1) CPU_SET(a, &set);
        if (comma == std::string::npos) break;
        pos = comma + 1;
    }
    if (sched_setaffinity(0, sizeof set, &set) != 0) return -1;
    snprintf(note, note_len, "node %d cpus %s", node, list.c_str());
    return node;
}

int main(int argc, char** argv) {
    int device = 0, reps = 5;
    double h2d_mb = 526, d2h_mb = 3146, start_at = 0;
    bool numa = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--h2d-mb") && i + 1 < argc) h2d_mb = atof(argv[++i]);
        else if (!strcmp(argv[i], "--d2h-mb") && i + 1 < argc) d2h_mb = atof(argv[++i]);
        else if (!strcmp(argv[i], "--start-at") && i + 1 < argc) start_at = atof(argv[++i]);
        else if (!strcmp(argv[i], "--numa")) numa = true;
    }
    CK(cudaSetDevice(device));
    char note[256] = "not bound";
    if (numa) bind_numa(device, note, sizeof note);
    const size_t nh = (size_t)(h2d_mb * 1e6), nd = (size_t)(d2h_mb * 1e6);
    uint8_t *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc((void**)&h_in, nh, cudaHostAllocDefault));
    CK(cudaHostAlloc((void**)&h_out, nd, cudaHostAllocDefault));
    memset(h_in, 1, nh);
    memset(h_out, 2, nd);
    CK(cudaMalloc((void**)&d_in, nh));
    CK(cudaMalloc((void**)&d_out, nd));
    CK(cudaMemset(d_out, 3, nd));
    cudaStream_t s_up, s_dn;
    CK(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_dn, cudaStreamNonBlocking));
    // warm-up
    CK(cudaMemcpyAsync(d_in, h_in, nh, cudaMemcpyHostToDevice, s_up));
    CK(cudaMemcpyAsync(h_out, d_out, nd, cudaMemcpyDeviceToHost, s_dn));
    CK(cudaDeviceSynchronize());
    while (start_at > 0 && now_s() < start_at) usleep(200);
    auto timed = [&](bool up, bool dn) -> double {
        double best = 1e30;
        for (int r = 0; r < reps; r++) {
            cudaDeviceSynchronize();
            const double t0 = now_s();
            if (up) cudaMemcpyAsync(d_in, h_in, nh, cudaMemcpyHostToDevice, s_up);
            if (dn) cudaMemcpyAsync(h_out, d_out, nd, cudaMemcpyDeviceToHost, s_dn);
            cudaDeviceSynchronize();
            const double dt = now_s() - t0;
            if (dt < best) best = dt;
        }
        return best;
    };
    // mean, not best, for the concurrent figure: ranks contend, and the step time a caller sees is the contended one
    auto timed_mean = [&](bool up, bool dn) -> double {
        cudaDeviceSynchronize();
        const double t0 = now_s();
        for (int r = 0; r < reps; r++) {
            if (up) cudaMemcpyAsync(d_in, h_in, nh, cudaMemcpyHostToDevice, s_up);
            if (dn) cudaMemcpyAsync(h_out, d_out, nd, cudaMemcpyDeviceToHost, s_dn);
            cudaDeviceSynchronize();
        }
        return (now_s() - t0) / reps;
    };
    const double t_both_mean = timed_mean(true, true);
    const double t_up = timed(true, false), t_dn = timed(false, true), t_both = timed(true, true);
    printf("{\"device\": %d, \"numa\": \"%s\", \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"reps\": %d, "
           "\"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"both_ms_best\": %.3f, \"both_ms_mean\": %.3f, \"h2d_ms\": %.3f, \"d2h_ms\": %.3f, "
           "\"both_gbs_sum\": %.2f}\n",
           device, note, nh, nd, reps, nh / t_up / 1e9, nd / t_dn / 1e9, t_both * 1e3, t_both_mean * 1e3, t_up * 1e3, t_dn * 1e3,
           (nh + nd) / t_both_mean / 1e9);
    return 0;
}
