// TEST HELPER (not product code): compiles fabric-mod_b200/csrc/ecdsa_verify.cuh for the host so the exact
// point / window / verify logic the CUDA kernel runs can be checked against the oracle on the GPU-less build
// box.  Only the innermost limb primitives differ between the two builds (PTX on device, uint64_t here); those
// are checked on the GPU by tests/test_gpu_field.py.
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include <utility>
#include "../../fabric-mod_b200/csrc/ecdsa_verify.cuh"

using namespace fabgpu;

static std::vector<aff> g_tab;

extern "C" {

void hostsim_build_gtable()
{
    if (!g_tab.empty()) return;
    g_tab.resize((size_t)FAB_G_WINDOWS * FAB_G_ENTRIES);
    // window j's entries from window j-1's by FAB_WG doublings would be faster; this mirrors the device kernel 1:1
    for (int j = 0; j < FAB_G_WINDOWS; j++)
        for (uint32_t d = 1; d <= (uint32_t)FAB_G_ENTRIES; d++)
            g_tab[(size_t)j * FAB_G_ENTRIES + d - 1] = g_table_entry(j, d);
}

// raw table bytes (little-endian limbs, Montgomery form): entries x 64 bytes
size_t hostsim_gtable(const uint8_t** out)
{
    hostsim_build_gtable();
    *out = (const uint8_t*)g_tab.data();
    return g_tab.size() * sizeof(aff);
}

// inputs: n x 32 big-endian bytes each; out[i] = V_INVALID / V_VALID / V_OFFCURVE
void hostsim_verify_batch(const uint8_t* qx, const uint8_t* qy, const uint8_t* e, const uint8_t* r, const uint8_t* s, int n, uint8_t* out)
{
    hostsim_build_gtable();
    for (int i = 0; i < n; i++) {
        size_t o = 32 * (size_t)i;
        out[i] = (uint8_t)ecdsa_verify_one(u256_from_be(qx + o), u256_from_be(qy + o), u256_from_be(e + o),
                                           u256_from_be(r + o), u256_from_be(s + o), g_tab.data());
    }
}

// Same batch through the per-key-table path: one table per distinct key (built with build_key_window, as the device's
// build_key_tables_kernel does), then ecdsa_verify_one_cached.  Keys that are not curve points are reported V_OFFCURVE.
void hostsim_verify_batch_cached(const uint8_t* qx, const uint8_t* qy, const uint8_t* e, const uint8_t* r, const uint8_t* s, int n, uint8_t* out)
{
    hostsim_build_gtable();
    std::vector<std::pair<std::vector<uint8_t>, std::vector<aff>>> cache;
    std::vector<u256> zs(2 * FAB_Q_ENTRIES);
    for (int i = 0; i < n; i++) {
        size_t o = 32 * (size_t)i;
        std::vector<uint8_t> key(qx + o, qx + o + 32);
        key.insert(key.end(), qy + o, qy + o + 32);
        const std::vector<aff>* tab = nullptr;
        for (auto& c : cache) if (c.first == key) tab = &c.second;
        if (!tab) {
            const u256 x = u256_from_be(qx + o), y = u256_from_be(qy + o);
            aff q; bool ok = u256_lt(x, fe_p()) && u256_lt(y, fe_p());
            if (ok) { q.x = fe_to_mont(x); q.y = fe_to_mont(y); ok = aff_on_curve(q); }
            std::vector<aff> t;
            if (ok) {
                t.resize((size_t)FAB_Q_WINDOWS * FAB_Q_ENTRIES);
                if (FAB_Q_TWO_LEVEL) {      // same two-level build as small_tables_kernel + full_tables_kernel
                    const int half = FAB_WQ / 2, nsmall = (1 << half) - 1;
                    std::vector<aff> lo(nsmall), hi(nsmall);
                    for (int j = 0; j < FAB_Q_WINDOWS; j++) {
                        build_multiples(q, FAB_WQ * j, nsmall, lo.data(), zs.data(), zs.data() + nsmall);
                        build_multiples(q, FAB_WQ * j + half, nsmall, hi.data(), zs.data(), zs.data() + nsmall);
                        for (uint32_t x0 = 1; x0 <= (uint32_t)FAB_Q_ENTRIES; x0 += FAB_TAB_CHUNK) {
                            const int cnt = (int)std::min<uint32_t>(FAB_TAB_CHUNK, (uint32_t)FAB_Q_ENTRIES - x0 + 1);
                            build_window_chunk(lo.data(), hi.data(), half, x0, cnt, t.data() + (size_t)j * FAB_Q_ENTRIES);
                        }
                    }
                } else
                for (int j = 0; j < FAB_Q_WINDOWS; j++) build_key_window(q, j, t.data() + (size_t)j * FAB_Q_ENTRIES, zs.data(), zs.data() + FAB_Q_ENTRIES);
            }
            cache.emplace_back(key, std::move(t));
            tab = &cache.back().second;
        }
        if (tab->empty()) { out[i] = (uint8_t)V_OFFCURVE; continue; }
        out[i] = (uint8_t)ecdsa_verify_one_cached(tab->data(), u256_from_be(e + o), u256_from_be(r + o), u256_from_be(s + o), g_tab.data());
    }
}

// field / scalar unit hooks: op 0 fe_mul, 1 fe_add, 2 fe_sub, 3 sc_mul, 4 fe_inv, 5 sc_inv_to_mont (b ignored for 4,5)
void hostsim_fieldop(int op, const uint8_t* a, const uint8_t* b, int n, uint8_t* out)
{
    for (int i = 0; i < n; i++) {
        u256 x = u256_from_be(a + 32 * (size_t)i), y = u256_from_be(b + 32 * (size_t)i), z;
        switch (op) {
            case 0: z = fe_mul(x, y); break;
            case 1: z = fe_add(x, y); break;
            case 2: z = fe_sub(x, y); break;
            case 3: z = sc_mul(x, y); break;
            case 4: z = fe_inv(x); break;
            case 5: z = sc_inv_to_mont(x); break;
        case 7: z = fe_sqr(x); break;
        default: z = sc_inv_to_mont_safegcd(x); break;
        }
        u256_to_be(z, out + 32 * (size_t)i);
    }
}

// k*P for affine plain (x,y) and plain scalar k, via the same table + Booth routine the kernel uses; returns affine plain
// coordinates, or all-zero for infinity.
void hostsim_scalar_mul(const uint8_t* k, const uint8_t* px, const uint8_t* py, uint8_t* ox, uint8_t* oy)
{
    aff q; q.x = fe_to_mont(u256_from_be(px)); q.y = fe_to_mont(u256_from_be(py));
    jac tab[16];
    build_q_table(tab, q);
    jac r = scalar_mul_var(u256_from_be(k), tab);
    if (jac_is_infinity(r)) { memset(ox, 0, 32); memset(oy, 0, 32); return; }
    aff a = jac_to_aff(r);
    u256_to_be(fe_from_mont(a.x), ox); u256_to_be(fe_from_mont(a.y), oy);
}

// k*G via the fixed-base table
void hostsim_base_mul(const uint8_t* k, uint8_t* ox, uint8_t* oy)
{
    hostsim_build_gtable();
    jac r = add_fixed_base(jac_infinity(), u256_from_be(k), g_tab.data());
    if (jac_is_infinity(r)) { memset(ox, 0, 32); memset(oy, 0, 32); return; }
    aff a = jac_to_aff(r);
    u256_to_be(fe_from_mont(a.x), ox); u256_to_be(fe_from_mont(a.y), oy);
}

}  // extern "C"
