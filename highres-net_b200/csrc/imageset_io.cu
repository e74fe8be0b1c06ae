// The I/O ends of the path, host side, no GPU involved:
//   * 16-bit greyscale PNG decode into caller-owned (pinned) uint16 buffers, many files at once on a thread pool --
//     replaces the `io.imread(join(imset_dir, f'LR{i}.png'))` loop of DataLoader.py:134-140 (skimage -> imageio -> PIL);
//   * clearance scores (sum of each QM status map, save_clearance.py:13-27) and the clearance order of the views
//     (DataLoader.py:118-131);
//   * 16-bit greyscale PNG encode + the stored ZIP archive of generate_submission_file (predict.py:161-194: io.imsave of
//     the img_as_uint image, then ZipFile(mode='w'), whose default is ZIP_STORED).
// PNG (ISO/IEC 15948) is restated from the published format: chunk layout, CRC-32, zlib stream in the IDAT chunks (zlib
// itself is the system library, as it is for PIL), the five scanline filters.  Non-interlaced greyscale only (colour type
// 0, bit depth 1/2/4/8/16): that is what the Proba-V files are.
#include "../../include/hrn_b200.h"
#include "internal.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hrn {
namespace {

struct PngHeader {
    uint32_t width = 0, height = 0;
    int bit_depth = 0, color_type = 0, interlace = 0;
};

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
void put_be32(uint8_t* p, uint32_t v) {
    p[0] = uint8_t(v >> 24), p[1] = uint8_t(v >> 16), p[2] = uint8_t(v >> 8), p[3] = uint8_t(v);
}

bool read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (f == nullptr) {
        err = std::string("cannot open ") + path;
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? static_cast<size_t>(n) : 0);
    const size_t got = out.empty() ? 0 : std::fread(out.data(), 1, out.size(), f);
    std::fclose(f);
    if (got != out.size()) {
        err = std::string("short read on ") + path;
        return false;
    }
    return true;
}

const uint8_t PNG_SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

// Walks the chunks: fills the header and appends the IDAT payloads (the zlib stream) to `idat` when it is non-null.
bool parse_png(const std::vector<uint8_t>& file, PngHeader& h, std::vector<uint8_t>* idat, std::string& err) {
    if (file.size() < 8 + 25 || std::memcmp(file.data(), PNG_SIG, 8) != 0) {
        err = "not a PNG file";
        return false;
    }
    size_t pos = 8;
    bool have_ihdr = false, have_iend = false;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + static_cast<size_t>(len) > file.size()) {
            err = "truncated PNG chunk";
            return false;
        }
        const uint8_t* data = &file[pos + 8];
        const uint32_t want_crc = be32(data + len);
        const uint32_t crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), type, len + 4));
        if (crc != want_crc) {
            err = "PNG chunk CRC mismatch";
            return false;
        }
        if (std::memcmp(type, "IHDR", 4) == 0) {
            if (len != 13) {
                err = "bad IHDR length";
                return false;
            }
            h.width = be32(data);
            h.height = be32(data + 4);
            h.bit_depth = data[8];
            h.color_type = data[9];
            h.interlace = data[12];
            if (data[10] != 0 || data[11] != 0) {
                err = "unknown PNG compression / filter method";
                return false;
            }
            have_ihdr = true;
            if (idat == nullptr) return true;
        } else if (std::memcmp(type, "IDAT", 4) == 0) {
            if (idat != nullptr) idat->insert(idat->end(), data, data + len);
        } else if (std::memcmp(type, "IEND", 4) == 0) {
            have_iend = true;
            break;
        } else if ((type[0] & 0x20) == 0) {      // an unknown CRITICAL chunk (PLTE included: palette images are rejected below)
            if (std::memcmp(type, "PLTE", 4) != 0) {
                err = "unknown critical PNG chunk";
                return false;
            }
        }
        pos += 12 + static_cast<size_t>(len);
    }
    if (!have_ihdr || !have_iend) {
        err = "PNG without IHDR / IEND";
        return false;
    }
    return true;
}

inline uint8_t paeth(uint8_t a, uint8_t b, uint8_t c) {
    const int p = int(a) + int(b) - int(c);
    const int pa = std::abs(p - int(a)), pb = std::abs(p - int(b)), pc = std::abs(p - int(c));
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Decodes one greyscale PNG into dst (height x width uint16, sample values as stored: 0/1 for 1-bit, 0..255 for 8-bit,
// 0..65535 for 16-bit -- what numpy gets from io.imread(...) cast to uint16).
bool decode_gray(const char* path, uint16_t* dst, int want_h, int want_w, std::string& err) {
    std::vector<uint8_t> file, z;
    if (!read_file(path, file, err)) return false;
    PngHeader h;
    if (!parse_png(file, h, &z, err)) {
        err += std::string(" (") + path + ")";
        return false;
    }
    if (h.color_type != 0 || h.interlace != 0 ||
        !(h.bit_depth == 1 || h.bit_depth == 2 || h.bit_depth == 4 || h.bit_depth == 8 || h.bit_depth == 16)) {
        err = std::string("unsupported PNG (need non-interlaced greyscale, got colour type ") + std::to_string(h.color_type) +
              ", bit depth " + std::to_string(h.bit_depth) + ", interlace " + std::to_string(h.interlace) + "): " + path;
        return false;
    }
    if (static_cast<int>(h.height) != want_h || static_cast<int>(h.width) != want_w) {
        err = std::string("PNG is ") + std::to_string(h.height) + " x " + std::to_string(h.width) + ", expected " +
              std::to_string(want_h) + " x " + std::to_string(want_w) + ": " + path;
        return false;
    }
    const size_t stride = (static_cast<size_t>(h.width) * h.bit_depth + 7) / 8;      // bytes per scanline, without the filter byte
    const size_t bpp = h.bit_depth == 16 ? 2 : 1;                                      // filter distance in bytes
    std::vector<uint8_t> raw((stride + 1) * h.height);
    uLongf raw_len = static_cast<uLongf>(raw.size());
    const int zr = uncompress(raw.data(), &raw_len, z.data(), static_cast<uLong>(z.size()));
    if (zr != Z_OK || raw_len != raw.size()) {
        err = std::string("PNG zlib stream is corrupt or has the wrong length: ") + path;
        return false;
    }
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (uint32_t y = 0; y < h.height; ++y) {
        uint8_t* line = &raw[(stride + 1) * y];
        const int filter = line[0];
        uint8_t* cur = line + 1;
        switch (filter) {
            case 0: break;
            case 1:
                for (size_t i = bpp; i < stride; ++i) cur[i] = uint8_t(cur[i] + cur[i - bpp]);
                break;
            case 2:
                for (size_t i = 0; i < stride; ++i) cur[i] = uint8_t(cur[i] + prev[i]);
                break;
            case 3:
                for (size_t i = 0; i < stride; ++i) {
                    const int left = i >= bpp ? cur[i - bpp] : 0;
                    cur[i] = uint8_t(cur[i] + ((left + int(prev[i])) >> 1));
                }
                break;
            case 4:
                for (size_t i = 0; i < stride; ++i) {
                    const uint8_t left = i >= bpp ? cur[i - bpp] : 0, up_left = i >= bpp ? prev[i - bpp] : 0;
                    cur[i] = uint8_t(cur[i] + paeth(left, prev[i], up_left));
                }
                break;
            default:
                err = std::string("unknown PNG filter type in ") + path;
                return false;
        }
        uint16_t* out = dst + static_cast<size_t>(y) * h.width;
        if (h.bit_depth == 16) {
            for (uint32_t x = 0; x < h.width; ++x) out[x] = uint16_t((cur[2 * x] << 8) | cur[2 * x + 1]);
        } else if (h.bit_depth == 8) {
            for (uint32_t x = 0; x < h.width; ++x) out[x] = cur[x];
        } else {
            const int per = 8 / h.bit_depth, mask = (1 << h.bit_depth) - 1;
            for (uint32_t x = 0; x < h.width; ++x)
                out[x] = uint16_t((cur[x / per] >> (8 - h.bit_depth * (x % per + 1))) & mask);
        }
        prev = cur;
    }
    return true;
}

void append_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, uint32_t len) {
    const size_t at = out.size();
    out.resize(at + 12 + len);
    put_be32(&out[at], len);
    std::memcpy(&out[at + 4], type, 4);
    if (len > 0) std::memcpy(&out[at + 8], data, len);
    put_be32(&out[at + 8 + len], static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), &out[at + 4], len + 4)));
}

// 16-bit greyscale, non-interlaced; every scanline uses the Up filter (rows of a natural image resemble the row above),
// deflate with the Z_RLE strategy: the filtered bytes of a 16-bit sensor image are noise-like, where LZ77 matching finds
// nothing and costs 3-5x the time (3.2 ms instead of 16 ms for a 384 x 384 image, and a slightly SMALLER file than level 6).
// Any conforming reader (PIL / skimage included) gets the same pixels back whatever the filter and strategy.
bool encode_gray16(const char* path, const uint16_t* src, int height, int width, std::string& err) {
    const size_t stride = static_cast<size_t>(width) * 2;
    std::vector<uint8_t> raw((stride + 1) * height);
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (int y = 0; y < height; ++y) {
        const uint16_t* row = src + static_cast<size_t>(y) * width;
        for (int x = 0; x < width; ++x) {
            cur[2 * x] = uint8_t(row[x] >> 8);
            cur[2 * x + 1] = uint8_t(row[x] & 0xff);
        }
        uint8_t* line = &raw[(stride + 1) * y];
        line[0] = 2;
        for (size_t i = 0; i < stride; ++i) line[1 + i] = uint8_t(cur[i] - prev[i]);
        prev.swap(cur);
    }
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, 6, Z_DEFLATED, 15, 9, Z_RLE) != Z_OK) {
        err = "zlib deflateInit2 failed";
        return false;
    }
    uLongf zlen = deflateBound(&zs, static_cast<uLong>(raw.size()));
    std::vector<uint8_t> z(zlen);
    zs.next_in = raw.data();
    zs.avail_in = static_cast<uInt>(raw.size());
    zs.next_out = z.data();
    zs.avail_out = static_cast<uInt>(z.size());
    const int zrc = deflate(&zs, Z_FINISH);
    zlen = zs.total_out;
    deflateEnd(&zs);
    if (zrc != Z_STREAM_END) {
        err = "zlib deflate failed";
        return false;
    }
    std::vector<uint8_t> out(PNG_SIG, PNG_SIG + 8);
    uint8_t ihdr[13];
    put_be32(ihdr, static_cast<uint32_t>(width));
    put_be32(ihdr + 4, static_cast<uint32_t>(height));
    ihdr[8] = 16, ihdr[9] = 0, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
    append_chunk(out, "IHDR", ihdr, 13);
    append_chunk(out, "IDAT", z.data(), static_cast<uint32_t>(zlen));
    append_chunk(out, "IEND", nullptr, 0);
    FILE* f = std::fopen(path, "wb");
    if (f == nullptr) {
        err = std::string("cannot create ") + path;
        return false;
    }
    const size_t put = std::fwrite(out.data(), 1, out.size(), f);
    const bool ok = std::fclose(f) == 0 && put == out.size();
    if (!ok) err = std::string("short write on ") + path;
    return ok;
}

// Runs fn(i) for i in [0, n) on `threads` host threads (0 = hardware concurrency); the first failure's message wins.
template <typename F>
int parallel_for(int n, int threads, F&& fn) {
    if (threads <= 0) threads = static_cast<int>(std::thread::hardware_concurrency());
    threads = std::max(1, std::min(threads, n));
    std::atomic<int> next{0}, failed{0};
    std::string first_error;
    std::mutex mu;
    auto work = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
            if (failed.load()) return;
            std::string err;
            if (!fn(i, err)) {
                std::lock_guard<std::mutex> lock(mu);
                if (!failed.exchange(1)) first_error = err;
            }
        }
    };
    if (threads == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    if (failed.load()) {
        set_error("%s", first_error.c_str());
        return -1;
    }
    return 0;
}

}  // namespace
}  // namespace hrn

extern "C" {

int32_t hrn_png_info(const char* path, int32_t* width, int32_t* height, int32_t* bit_depth, int32_t* color_type) {
    if (path == nullptr || width == nullptr || height == nullptr || bit_depth == nullptr || color_type == nullptr) {
        hrn::set_error("hrn_png_info: null argument");
        return -1;
    }
    std::vector<uint8_t> file;
    std::string err;
    hrn::PngHeader h;
    if (!hrn::read_file(path, file, err) || !hrn::parse_png(file, h, nullptr, err)) {
        hrn::set_error("hrn_png_info: %s", err.c_str());
        return -1;
    }
    *width = static_cast<int32_t>(h.width);
    *height = static_cast<int32_t>(h.height);
    *bit_depth = h.bit_depth;
    *color_type = h.color_type;
    return 0;
}

int32_t hrn_png_read_gray_u16(const char* const* paths, int32_t n, int32_t height, int32_t width, uint16_t* dst,
                              int32_t threads) {
    if (paths == nullptr || dst == nullptr || n < 0 || height <= 0 || width <= 0) {
        hrn::set_error("hrn_png_read_gray_u16: bad argument");
        return -1;
    }
    const size_t plane = static_cast<size_t>(height) * width;
    return hrn::parallel_for(n, threads, [&](int i, std::string& err) {
        if (paths[i] == nullptr) {
            err = "hrn_png_read_gray_u16: null path";
            return false;
        }
        return hrn::decode_gray(paths[i], dst + plane * i, height, width, err);
    });
}

int32_t hrn_png_write_gray_u16(const char* const* paths, int32_t n, int32_t height, int32_t width, const uint16_t* src,
                               int32_t threads) {
    if (paths == nullptr || src == nullptr || n < 0 || height <= 0 || width <= 0) {
        hrn::set_error("hrn_png_write_gray_u16: bad argument");
        return -1;
    }
    const size_t plane = static_cast<size_t>(height) * width;
    return hrn::parallel_for(n, threads, [&](int i, std::string& err) {
        if (paths[i] == nullptr) {
            err = "hrn_png_write_gray_u16: null path";
            return false;
        }
        return hrn::encode_gray16(paths[i], src + plane * i, height, width, err);
    });
}

int32_t hrn_clearance_scores(const char* const* qm_paths, int32_t n, int32_t height, int32_t width, int32_t threads,
                             double* scores) {
    if (qm_paths == nullptr || scores == nullptr || n < 0 || height <= 0 || width <= 0) {
        hrn::set_error("hrn_clearance_scores: bad argument");
        return -1;
    }
    const size_t plane = static_cast<size_t>(height) * width;
    return hrn::parallel_for(n, threads, [&](int i, std::string& err) {
        std::vector<uint16_t> map(plane);
        if (qm_paths[i] == nullptr || !hrn::decode_gray(qm_paths[i], map.data(), height, width, err)) {
            if (err.empty()) err = "hrn_clearance_scores: null path";
            return false;
        }
        unsigned long long sum = 0;                       // lr_maps.sum(axis=(1, 2)) of the uint16 array (numpy sums in uint64)
        for (size_t k = 0; k < plane; ++k) sum += map[k];
        scores[i] = static_cast<double>(sum);
        return true;
    });
}

int32_t hrn_clearance_order(const double* clearances, int32_t n, int32_t* order) {
    if (clearances == nullptr || order == nullptr || n < 0) {
        hrn::set_error("hrn_clearance_order: bad argument");
        return -1;
    }
    // np.argsort(clearances)[::-1]: ascending, then reversed.  Ties: numpy's default sort is not stable for n > 16, so the
    // reference's order of equally clear views is implementation-defined there; here it is the reversed STABLE order
    // (equal scores by descending index), which is what numpy itself yields for n <= 16.
    std::vector<int32_t> idx(n);
    for (int32_t i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return clearances[a] < clearances[b]; });
    for (int32_t i = 0; i < n; ++i) order[i] = idx[n - 1 - i];
    return 0;
}

int32_t hrn_zip_store(const char* zip_path, const char* const* files, const char* const* arcnames, int32_t n) {
    if (zip_path == nullptr || files == nullptr || arcnames == nullptr || n < 0 || n > 65535) {
        hrn::set_error("hrn_zip_store: bad argument (at most 65535 members)");
        return -1;
    }
    FILE* f = std::fopen(zip_path, "wb");
    if (f == nullptr) {
        hrn::set_error("hrn_zip_store: cannot create %s", zip_path);
        return -1;
    }
    auto le16 = [](std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x)), v.push_back(uint8_t(x >> 8)); };
    auto le32 = [&](std::vector<uint8_t>& v, uint32_t x) { le16(v, x & 0xffff), le16(v, x >> 16); };
    std::vector<uint8_t> central;
    uint64_t offset = 0;
    bool ok = true;
    for (int32_t i = 0; i < n && ok; ++i) {
        std::vector<uint8_t> data;
        std::string err;
        if (files[i] == nullptr || arcnames[i] == nullptr || !hrn::read_file(files[i], data, err)) {
            hrn::set_error("hrn_zip_store: %s", err.empty() ? "null member" : err.c_str());
            ok = false;
            break;
        }
        const std::string name(arcnames[i]);
        const uint32_t crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), data.data(), static_cast<uInt>(data.size())));
        if (data.size() > 0xfffffffeull || offset > 0xfffffffeull) {
            hrn::set_error("hrn_zip_store: member too large for a plain (non-zip64) archive");
            ok = false;
            break;
        }
        std::vector<uint8_t> local;
        le32(local, 0x04034b50), le16(local, 20), le16(local, 0), le16(local, 0);       // version 2.0, no flags, method 0 = stored
        le16(local, 0), le16(local, 0x21);                                             // time 00:00:00, date 1980-01-01
        le32(local, crc), le32(local, static_cast<uint32_t>(data.size())), le32(local, static_cast<uint32_t>(data.size()));
        le16(local, static_cast<uint32_t>(name.size())), le16(local, 0);
        local.insert(local.end(), name.begin(), name.end());
        ok = std::fwrite(local.data(), 1, local.size(), f) == local.size() &&
             (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size());
        le32(central, 0x02014b50), le16(central, 20), le16(central, 20), le16(central, 0), le16(central, 0);
        le16(central, 0), le16(central, 0x21);
        le32(central, crc), le32(central, static_cast<uint32_t>(data.size())), le32(central, static_cast<uint32_t>(data.size()));
        le16(central, static_cast<uint32_t>(name.size())), le16(central, 0), le16(central, 0), le16(central, 0), le16(central, 0);
        le32(central, 0), le32(central, static_cast<uint32_t>(offset));
        central.insert(central.end(), name.begin(), name.end());
        offset += local.size() + data.size();
    }
    if (ok) {
        std::vector<uint8_t> end;
        le32(end, 0x06054b50), le16(end, 0), le16(end, 0), le16(end, static_cast<uint32_t>(n)), le16(end, static_cast<uint32_t>(n));
        le32(end, static_cast<uint32_t>(central.size())), le32(end, static_cast<uint32_t>(offset)), le16(end, 0);
        ok = (central.empty() || std::fwrite(central.data(), 1, central.size(), f) == central.size()) &&
             std::fwrite(end.data(), 1, end.size(), f) == end.size();
        if (!ok) hrn::set_error("hrn_zip_store: short write on %s", zip_path);
    }
    if (std::fclose(f) != 0 && ok) {
        hrn::set_error("hrn_zip_store: close failed on %s", zip_path);
        ok = false;
    }
    return ok ? 0 : -1;
}

}  // extern "C"
