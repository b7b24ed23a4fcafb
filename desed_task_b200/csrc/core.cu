// Error plumbing and small queries of libsedk.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <ctype.h>
#include <map>
#include <string>
#include <vector>

namespace sedk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void count_launch() { g_launches++; }

// ---- named integer options (kernel-variant switches for A/B measurements and parity tests); default from the environment
// variable SEDK_<NAME> (upper case) on first use, else the caller's default
static std::map<std::string, int> g_opts;
int get_option(const char* name, int dflt) {
    auto it = g_opts.find(name);
    if (it != g_opts.end()) return it->second;
    std::string env = "SEDK_";
    for (const char* c = name; *c; c++) env += (char)toupper(*c);
    const char* e = getenv(env.c_str());
    int v = (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
    g_opts[name] = v;
    return v;
}
void set_option(const char* name, int value) { g_opts[name] = value; }

// ---- optional per-launcher device timing (eager mode only; never enabled inside a timed benchmark region)
struct ProfEntry { std::string name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;

ProfScope::ProfScope(const char* name, cudaStream_t s) : idx_(-1), s_(s) {
    if (!g_prof_on) return;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
    ProfEntry e;
    e.name = name;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, s);
    g_prof.push_back(e);
    idx_ = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx_ >= 0) cudaEventRecord(g_prof[idx_].b, s_);
}

bool profiling_on() { return g_prof_on; }

int check_launch(const char* what) {
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SEDK_ERR_CUDA;
    }
    return SEDK_OK;
}

}  // namespace sedk

extern "C" const char* sedk_last_error(void) { return sedk::g_err; }
extern "C" int sedk_version(void) { return 100; }
extern "C" int sedk_device_cc(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
    return major * 10 + minor;
}
extern "C" int sedk_sizeof_crnn_plan(void) { return (int)sizeof(sedk_crnn_plan); }
extern "C" long long sedk_launch_count(void) { return sedk::g_launches; }

extern "C" int sedk_set_option(const char* name, int value) {
    if (name == nullptr) return SEDK_ERR_INVALID;
    sedk::set_option(name, value);
    return SEDK_OK;
}
extern "C" int sedk_get_option(const char* name, int dflt) { return name ? sedk::get_option(name, dflt) : dflt; }

extern "C" int sedk_profile_enable(int on) {
    using namespace sedk;
    for (auto& e : g_prof) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof.clear();
    g_prof_on = on != 0;
    return SEDK_OK;
}

extern "C" int sedk_profile_report(char* buf, int buflen) {
    using namespace sedk;
    if (buf == nullptr || buflen <= 0) return SEDK_ERR_INVALID;
    if (cudaDeviceSynchronize() != cudaSuccess) return SEDK_ERR_CUDA;
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& e : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) continue;
        auto& a = agg[e.name];
        a.first += 1;
        a.second += ms;
    }
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if ((int)out.size() + 1 > buflen) return SEDK_ERR_INVALID;
    memcpy(buf, out.c_str(), out.size() + 1);
    return SEDK_OK;
}
