// TEST INFRASTRUCTURE.  Link-time replacement of libc's rand()/srand() with a thread-local generator, for the third CPU
// figure SURVEY.md section 8(d) asks for: "N threads with a thread-local rand() interposed at link time (no reference
// source edit)".  glibc's rand() takes a process-wide lock, so the reference's N render threads serialise on it in every
// stochastic branch (lens, soft shadows, glossy lobes, MonteCarlo).  An executable's own definition of rand takes
// precedence over libc's for every object linked into it, the reference's included.
#include <atomic>
#include <cstdint>
#include <cstdlib>

static std::atomic<uint64_t> g_seed{1};
static std::atomic<uint64_t> g_thread{0};

static uint64_t &state()
{
    thread_local uint64_t s = 0;
    thread_local uint64_t seen = ~0ull;
    const uint64_t seed = g_seed.load(std::memory_order_relaxed);
    if (seen != seed) { // (re)seed this thread's stream: splitmix of (seed, thread ordinal)
        seen = seed;
        uint64_t z = seed * 0x9E3779B97F4A7C15ull + (g_thread.fetch_add(1) + 1) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        s = (z ^ (z >> 31)) | 1ull;
    }
    return s;
}

extern "C" int rand(void)
{
    uint64_t &s = state();
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27; // xorshift64*
    return (int)(((s * 2685821657736338717ull) >> 33) & (uint64_t)RAND_MAX);
}

extern "C" void srand(unsigned seed) { g_seed.store((uint64_t)seed + 0x100000000ull * (g_seed.load() >> 32) + 0x100000000ull); }
