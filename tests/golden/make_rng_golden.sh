#!/usr/bin/env bash
# Generates tests/golden/rng_golden.bin from the REFERENCE's own RNG header
# (/root/reference/src/pcg_random.hpp, RNG = pcg32_k64_fast, src/commondef.h:63) and this
# container's libstdc++ distributions -- the exact objects the reference draws through
# (std::uniform_real_distribution<float>, std::normal_distribution<float>).
# Layout (little endian): for each seed in {0, 1, 12345, 4294967295+7}:
#   u32 raw[4096] | f32 uniform[4096] | f32 normal(0,1)[4096]   (each stream from a fresh RNG(seed))
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
cat > /tmp/rng_golden.cpp <<'CPP'
#include <cstdio>
#include <random>
#include <vector>
#include <cstdint>
#include "/root/reference/src/pcg_random.hpp"
typedef pcg32_k64_fast RNG;
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "wb");
    const uint64_t seeds[4] = {0ULL, 1ULL, 12345ULL, 4294967295ULL + 7ULL};
    for (uint64_t seed : seeds) {
        std::vector<uint32_t> raw(4096); std::vector<float> u(4096), n(4096);
        { RNG rng(seed); for (auto &x : raw) x = rng(); }
        { RNG rng(seed); std::uniform_real_distribution<float> d(0.0f, 1.0f); for (auto &x : u) x = d(rng); }
        { RNG rng(seed); std::normal_distribution<float> d(0.0f, 1.0f); for (auto &x : n) x = d(rng); }
        fwrite(raw.data(), 4, 4096, f); fwrite(u.data(), 4, 4096, f); fwrite(n.data(), 4, 4096, f);
    }
    fclose(f);
    return 0;
}
CPP
/usr/bin/g++ -O2 -std=c++11 -o /tmp/rng_golden /tmp/rng_golden.cpp
/tmp/rng_golden "$HERE/rng_golden.bin"
ls -la "$HERE/rng_golden.bin"
