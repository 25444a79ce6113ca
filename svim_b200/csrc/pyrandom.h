// Host restatement of the CPython 3.12 `random` pieces that SVIM_clustering.py:129-134
// depends on: `seed(1524)` (MT19937 init_by_array) and `sample(population, 100)`
// (Lib/random.py: pool algorithm for n <= 1045, set-based rejection above).
// The Mersenne-Twister stream is inherently sequential across the partitions of one
// type, so this runs on the host between two kernel launches; it consumes only the
// partition sizes.  Verified against the stdlib in tests/test_sampling.py.
#pragma once
#include <stdint.h>
#include <unordered_set>
#include <vector>

struct PyRandom {
    uint32_t mt[624];
    int idx;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void seed_int(uint32_t key) {   // random.seed(int) with a key that fits one 32-bit word
        init_genrand(19650218u);
        int i = 1, j = 0;
        const uint32_t keys[1] = {key};
        for (int k = 624; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + keys[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= 1) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
        idx = 624;
    }
    uint32_t next_u32() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    uint64_t randbelow(uint64_t n) {   // n < 2^32
        int k = 0;
        for (uint64_t t = n; t; t >>= 1) ++k;
        uint64_t r;
        do { r = next_u32() >> (32 - k); } while (r >= n);
        return r;
    }
    // indices chosen by random.sample(range(n), 100), n > 100
    void sample100(uint64_t n, int32_t* out) {
        const int K = 100;
        if (n <= 1045) {
            std::vector<int32_t> pool(n);
            for (uint64_t i = 0; i < n; ++i) pool[i] = (int32_t)i;
            for (int i = 0; i < K; ++i) {
                uint64_t j = randbelow(n - i);
                out[i] = pool[j];
                pool[j] = pool[n - i - 1];
            }
        } else {
            std::unordered_set<uint64_t> sel;
            for (int i = 0; i < K; ++i) {
                uint64_t j = randbelow(n);
                while (sel.count(j)) j = randbelow(n);
                sel.insert(j);
                out[i] = (int32_t)j;
            }
        }
    }
};
