// Row partition shared by the tcgen05 conv kernels: the flattened (image, column tile, row) space is split evenly
// over the CTA groups and walked as per-image strips.
#pragma once
#include "internal.h"

namespace hrn {

struct Geometry {
    int n_parts;      // cout / 64
    int groups;       // CTAs per part
    int x_tiles;      // ceil(W / 128)
    long long total_rows;   // n_img * x_tiles * H
    int split;        // ranges per CTA group (round-robin), >= 1
    int even_rows;    // 1: every CTA range starts and ends on an even row (the fused 2x2 max pool pairs rows 2k, 2k + 1)
    int img_group;    // images per M tile: 1 = 128 pixels of one image row; G > 1 = the rows of G narrow images side by
                      // side, (W + 2)-pixel segments (conv3x3_umma only; the "images" the walker sees are then groups)
};

struct Strip {
    int m, xt, y0, rows;
};

// The range of flattened rows owned by CTA group gi, cut into per-image strips.  With a live-work list the flattened
// space covers only the *live_count listed images (the count is known on the device only, so every CTA derives its
// own share from it) and Strip::m is looked up in the list; the MMA role never needs m and passes want_m = false.
struct StripWalker {
    long long total, g, g_end;
    int H, x_tiles, vg, vg_total, vg_stride;
    long long row_mask;              // ~1 when ranges must start on even rows (H is even then), ~0 otherwise
    const int* list;
    // The flattened space is cut into groups * split equal ranges dealt round-robin to the CTA groups (split = 1: one
    // contiguous range per CTA).
    __device__ StripWalker(const Geometry& geo, const ConvArgs& a, int gi, bool want_m = true)
        : H(a.H), x_tiles(geo.x_tiles), vg(gi), vg_total(geo.groups * geo.split), vg_stride(geo.groups),
          row_mask(geo.even_rows ? ~1LL : ~0LL), list(want_m ? a.live_list : nullptr) {
        total = geo.total_rows;
        if (a.live_count != nullptr) total = static_cast<long long>(*a.live_count) * geo.x_tiles * a.H;
        g = (total * vg / vg_total) & row_mask;
        g_end = (total * (vg + 1) / vg_total) & row_mask;
    }
    __device__ bool next(Strip& s) {
        while (g >= g_end) {
            vg += vg_stride;
            if (vg >= vg_total) return false;
            g = (total * vg / vg_total) & row_mask;
            g_end = (total * (vg + 1) / vg_total) & row_mask;
        }
        const long long col = g / H;
        s.y0 = static_cast<int>(g % H);
        s.rows = static_cast<int>(min(static_cast<long long>(H - s.y0), g_end - g));
        s.m = static_cast<int>(col / x_tiles);
        if (list != nullptr) s.m = list[s.m];
        s.xt = static_cast<int>(col % x_tiles);
        g += s.rows;
        return true;
    }
};

}  // namespace hrn
