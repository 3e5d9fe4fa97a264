// Per-device watchdog status block and launch configuration of the tensor-core kernels.
//
// Every tcgen05 kernel bounds its mbarrier waits (umma.cuh).  A wait that gives up records a code in the status word
// of the device it runs on -- pinned host memory mapped into the device's address space, so the host can read it with
// a plain load, without a CUDA call and without synchronising -- and traps.  One 64 KB block per device, allocated
// at the first tensor-core launch on that device and kept for the life of the process; nothing else is allocated.
// Word 0 is the status word; words 1024.. are a scratch area that only the timeline build of chain_split.cu
// (-DAL3D_SPLIT_TIMELINE, scripts/split_timeline.py) writes.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/al3d.h"

namespace al3d {

namespace {
constexpr int kMaxDevices = 64;
constexpr size_t kSlotBytes = 65536;
struct Slot { unsigned int *host = nullptr; unsigned int *dev = nullptr; int sms = 0; };
Slot g_slot[kMaxDevices];
std::mutex g_mu;
int g_trap = -1, g_stress = -1;          // -1: take the default from the environment at first use

int env_int(const char *name, int dflt)
{
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

int slot_for_current_device(Slot **out)
{
    int dev = 0;
    AL3D_CHECK_CUDA(cudaGetDevice(&dev));
    AL3D_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_mu);
    Slot &s = g_slot[dev];
    if (!s.host) {
        void *h = nullptr, *d = nullptr;
        AL3D_CHECK_CUDA(cudaHostAlloc(&h, kSlotBytes, cudaHostAllocMapped | cudaHostAllocPortable));
        std::memset(h, 0, kSlotBytes);
        AL3D_CHECK_CUDA(cudaHostGetDevicePointer(&d, h, 0));
        int sms = 0;
        AL3D_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        s.host = (unsigned int *)h; s.dev = (unsigned int *)d; s.sms = sms > 0 ? sms : 148;
    }
    if (g_trap < 0) g_trap = env_int("AL3D_TC_TRAP", 1) ? 1 : 0;
    if (g_stress < 0) g_stress = env_int("AL3D_TC_STRESS_NS", 0);
    *out = &s;
    return 0;
}
}  // namespace

int tc_launch_status(umma::TcStatus *out)
{
    Slot *s = nullptr;
    if (slot_for_current_device(&s)) return 1;
    out->word = s->dev; out->trap = g_trap; out->stress = g_stress;
    return 0;
}

int tc_num_sms()
{
    Slot *s = nullptr;
    if (slot_for_current_device(&s)) return 148;
    return s->sms;
}

}  // namespace al3d

using namespace al3d;

extern "C" int al3d_tc_abort_code(int *code_host)
{
    Slot *s = nullptr;
    if (slot_for_current_device(&s)) return 1;
    const unsigned int v = *(volatile unsigned int *)s->host;
    if (v != 0) *(volatile unsigned int *)s->host = 0;
    if (code_host) *code_host = (int)v;
    return 0;
}

extern "C" int al3d_tc_status_word_host(const void **word_host)
{
    AL3D_CHECK_ARG(word_host, "al3d_tc_status_word_host: null pointer");
    Slot *s = nullptr;
    if (slot_for_current_device(&s)) return 1;
    *word_host = s->host;
    return 0;
}

extern "C" int al3d_tc_configure(int trap_on_timeout, int stress_ns)
{
    AL3D_CHECK_ARG(stress_ns >= 0 && (stress_ns & 0xFFFFFF) <= 1000000 && (stress_ns >> 25) == 0, "al3d_tc_configure: stress_ns=%d", stress_ns);
    std::lock_guard<std::mutex> lock(g_mu);
    g_trap = trap_on_timeout & 3;         // bit 1: test hook, seg_pass1_kernel with the round-1 single release barrier
    g_stress = stress_ns;
    return 0;
}
