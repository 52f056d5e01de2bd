// Test-only host build of the bit-parallel column step in csrc/apm_myers.cuh (the __host__ path of
// myers_step / add_chain).  It lets the CPU test-suite check the recurrence, the carry chain across
// words and the masked final score against the oracle without a GPU.  Not part of the product.
#include <cstdint>
#include <cstring>

#include "../inf560-approximate-pattern-matching_b200/csrc/apm_myers.cuh"

template <int NW>
static int dist(const uint8_t *pat, const uint8_t *win, int len) {
    uint32_t peq[256][NW];
    std::memset(peq, 0, sizeof peq);
    for (int i = 0; i < len; ++i) peq[pat[i]][i >> 5] |= 1u << (i & 31);
    uint32_t Pv[NW], Mv[NW];
    for (int w = 0; w < NW; ++w) { Pv[w] = 0xFFFFFFFFu; Mv[w] = 0; }
    for (int c = 0; c < len; ++c) apm::myers_step<NW>(Pv, Mv, peq[win[c]]);
    const int topbits = len - 32 * (NW - 1);
    const uint32_t topmask = topbits >= 32 ? 0xFFFFFFFFu : ((1u << topbits) - 1u);
    return len + apm::myers_score_minus_len<NW>(Pv, Mv, topmask);
}

extern "C" int host_myers_distance(const uint8_t *pat, const uint8_t *win, int len) {
    if (len <= 0 || len > 256) return -1;
    switch ((len + 31) / 32) {
        case 1: return dist<1>(pat, win, len);
        case 2: return dist<2>(pat, win, len);
        case 3: return dist<3>(pat, win, len);
        case 4: return dist<4>(pat, win, len);
        case 5: return dist<5>(pat, win, len);
        case 6: return dist<6>(pat, win, len);
        case 7: return dist<7>(pat, win, len);
        default: return dist<8>(pat, win, len);
    }
}
