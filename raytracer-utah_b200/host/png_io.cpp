// PNG reader/writer on top of zlib.  The reference uses LodePNG for both directions
// (texture.cpp:75 lodepng::decode(...,LCT_RGB) and scene.h:644-654 lodepng::encode); the
// codec is third-party and off the hot path, so only the
// container is implemented here: 8/16-bit grey, grey+alpha, RGB, RGBA and palette images,
// non-interlaced, converted to 8-bit RGB like LCT_RGB/8 does (alpha dropped, 16-bit -> high byte).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host_scene.h"

namespace rtu {

namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(std::vector<uint8_t> &o, uint32_t v)
{
    o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v);
}
int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

} // namespace

bool decode_png_rgb8(const char *path, std::vector<uint8_t> *rgb, int *w, int *h, std::string *err)
{
    auto fail = [&](const char *m) { if (err) *err = std::string(path) + ": " + m; return false; };
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail("cannot open");
    std::vector<uint8_t> file;
    {
        uint8_t tmp[1 << 16];
        size_t n;
        while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) file.insert(file.end(), tmp, tmp + n);
        fclose(fp);
    }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) return fail("not a PNG");
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) return fail("truncated chunk");
        const uint8_t *body = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail("bad IHDR");
            W = be32(body); H = be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (W == 0 || H == 0) return fail("missing IHDR");
    // the header is untrusted input: the same limit frame_dims() applies to images (2^28 pixels), and 16-bit sides
    if (W > 65535u || H > 65535u || (uint64_t)W * H > (1ull << 28)) return fail("image dimensions out of range");
    if (interlace != 0) return fail("interlaced PNG not supported");
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) return fail("bad colour type");
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return fail("bad bit depth");
    size_t bpp_bits = (size_t)channels * depth;
    size_t stride = (W * bpp_bits + 7) / 8;
    size_t bpp = (bpp_bits + 7) / 8; // filter distance in bytes
    std::vector<uint8_t> raw((stride + 1) * (size_t)H);
    uLongf outlen = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &outlen, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || outlen != raw.size()) return fail("inflate failed");
    // unfilter in place
    std::vector<uint8_t> img(stride * (size_t)H);
    for (uint32_t y = 0; y < H; y++) {
        const uint8_t *src = &raw[(stride + 1) * (size_t)y];
        uint8_t *cur = &img[stride * (size_t)y];
        const uint8_t *up = y ? &img[stride * (size_t)(y - 1)] : nullptr;
        int ft = src[0];
        src++;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= bpp ? cur[x - bpp] : 0;
            int b = up ? up[x] : 0;
            int c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: return fail("bad filter type");
            }
            cur[x] = (uint8_t)v;
        }
    }
    rgb->resize((size_t)W * H * 3);
    for (uint32_t y = 0; y < H; y++) {
        const uint8_t *row = &img[stride * (size_t)y];
        for (uint32_t x = 0; x < W; x++) {
            uint8_t *o = &(*rgb)[((size_t)y * W + x) * 3];
            auto sample = [&](int ch) -> int { // 8-bit value of channel ch of pixel x
                if (depth == 8) return row[(size_t)x * channels + ch];
                if (depth == 16) return row[((size_t)x * channels + ch) * 2];
                size_t bit = (size_t)x * depth;
                int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                return v;
            };
            if (ctype == 2 || ctype == 6) { o[0] = sample(0); o[1] = sample(1); o[2] = sample(2); }
            else if (ctype == 0 || ctype == 4) {
                int g = sample(0);
                if (depth < 8) g = g * 255 / ((1 << depth) - 1);
                o[0] = o[1] = o[2] = (uint8_t)g;
            } else { // palette
                size_t idx = (size_t)sample(0);
                if (idx * 3 + 2 < plte.size()) { o[0] = plte[idx * 3]; o[1] = plte[idx * 3 + 1]; o[2] = plte[idx * 3 + 2]; }
                else { o[0] = o[1] = o[2] = 0; }
            }
        }
    }
    *w = (int)W;
    *h = (int)H;
    return true;
}

// LoadPPM (texture.cpp:18-55): binary P6; the header is read line by line (a line ends at \n, \r or after 1024 characters),
// '#' lines are skipped before the size line and before the maxval line, the pixels follow the maxval line directly.
// Like the reference, the magic is only rejected when BOTH of its characters are wrong and the maxval is not looked at.
bool decode_ppm_rgb8(const char *path, std::vector<uint8_t> *rgb, int *w, int *h, std::string *err)
{
    auto fail = [&](const char *m) { if (err) *err = std::string(path) + ": " + m; return false; };
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail("cannot open file");
    char buf[1024];
    auto read_line = [&]() {
        int i;
        for (i = 0; i < 1024; i++) {
            int c = fgetc(fp);
            buf[i] = (char)c;
            if (c == EOF || c == '\n' || c == '\r') { buf[i] = '\0'; return; }
        }
        buf[1023] = '\0';
    };
    read_line();
    if (buf[0] != 'P' && buf[1] != '6') { fclose(fp); return fail("not a PPM file"); }
    read_line();
    while (buf[0] == '#') read_line();
    int W = 0, H = 0;
    sscanf(buf, "%d %d", &W, &H);
    read_line();
    while (buf[0] == '#') read_line();
    if (W <= 0 || H <= 0 || W > 65535 || H > 65535 || (long long)W * H > (1ll << 28)) { fclose(fp); return fail("image dimensions out of range"); }
    rgb->assign((size_t)W * H * 3, 0);
    size_t got = fread(rgb->data(), 3, (size_t)W * H, fp);
    (void)got; // a short file leaves the rest of the pixels as they were allocated (zero here)
    fclose(fp);
    *w = W;
    *h = H;
    return true;
}

bool encode_png(const char *path, const uint8_t *px, int w, int h, int channels, std::string *err)
{
    auto fail = [&](const char *m) { if (err) *err = std::string(path) + ": " + m; return false; };
    if (w <= 0 || h <= 0 || (channels != 1 && channels != 3)) return fail("bad image");
    size_t stride = (size_t)w * channels;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    for (int y = 0; y < h; y++) {
        raw[(stride + 1) * (size_t)y] = 0; // filter: none
        memcpy(&raw[(stride + 1) * (size_t)y + 1], px + stride * (size_t)y, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail("deflate failed");
    comp.resize(clen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char *type, const std::vector<uint8_t> &body) {
        put32(out, (uint32_t)body.size());
        size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), body.begin(), body.end());
        uint32_t crc = (uint32_t)crc32(0L, &out[start], (uInt)(out.size() - start));
        put32(out, crc);
    };
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w);
    put32(ihdr, (uint32_t)h);
    ihdr.push_back(8);
    ihdr.push_back(channels == 3 ? 2 : 0); // LCT_RGB / LCT_GREY (scene.h:646-650)
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr);
    chunk("IDAT", comp);
    chunk("IEND", {});
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail("cannot create");
    size_t n = fwrite(out.data(), 1, out.size(), fp);
    fclose(fp);
    return n == out.size() ? true : fail("short write");
}

} // namespace rtu
