// tests/cpp/texconv_mini.cpp — a caller written the way a DirectXTex user writes one (cf. the call sites in the
// reference's Texconv/texconv.cpp:3109 Convert, :3434 GenerateMipMaps, :3707-3711 Compress), built against
// DirectXTexB200.h and linked to libdxtex_b200.so.  Used by tests/test_gpu_cpp_api.py.
//   texconv_mini <op> <in.raw> <out.raw> <width> <height> <srcfmt> <dstfmt|filter> [flags] [items]
#include "DirectXTexB200.h"
#include "../../include/dxtex_b200.h"      // dxb200_init_devices: the one call a multi-GPU caller adds
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace DirectX;

static std::vector<uint8_t> slurp(const char* p)
{
    std::vector<uint8_t> v; FILE* f = fopen(p, "rb"); if (!f) return v;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); v.resize(size_t(n));
    if (fread(v.data(), 1, size_t(n), f) != size_t(n)) v.clear();
    fclose(f); return v;
}
static bool dump(const char* p, const uint8_t* d, size_t n)
{
    FILE* f = fopen(p, "wb"); if (!f) return false;
    const bool ok = fwrite(d, 1, n, f) == n; fclose(f); return ok;
}

int main(int argc, char** argv)
{
    if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
    const std::string op = argv[1];
    auto in = slurp(argv[2]);
    const size_t w = strtoull(argv[4], nullptr, 10), h = strtoull(argv[5], nullptr, 10);
    const DXGI_FORMAT sf = DXGI_FORMAT(atoi(argv[6]));
    const uint32_t arg = uint32_t(strtoul(argv[7], nullptr, 0));
    const uint32_t flags = argc > 8 ? uint32_t(strtoul(argv[8], nullptr, 0)) : 0;
    const size_t items = argc > 9 ? strtoull(argv[9], nullptr, 10) : 1;

    size_t rowPitch = 0, slicePitch = 0;
    if (FAILED(ComputePitch(sf, w, h, rowPitch, slicePitch))) { fprintf(stderr, "pitch\n"); return 3; }
    if (in.size() < slicePitch * items) { fprintf(stderr, "short input\n"); return 3; }
    std::vector<Image> imgs(items);
    for (size_t i = 0; i < items; ++i) imgs[i] = Image{ w, h, sf, rowPitch, slicePitch, in.data() + i * slicePitch };
    TexMetadata md{}; md.width = w; md.height = h; md.depth = 1; md.arraySize = items; md.mipLevels = 1; md.format = sf; md.dimension = TEX_DIMENSION_TEXTURE2D;

    // TEXCONV_MINI_DEVICES=0,1,2,...: shard array calls / large images over these GPUs (dxb200_init_devices)
    if (const char* env = getenv("TEXCONV_MINI_DEVICES"))
    {
        std::vector<int> devs;
        for (const char* c = env; *c;) { devs.push_back(atoi(c)); while (*c && *c != ',') ++c; if (*c == ',') ++c; }
        const HRESULT hi = dxb200_init_devices(int(devs.size()), devs.data());
        int got[16]; const int n = dxb200_initialized_devices(got, 16);
        printf("devices=%d\n", n);
        if (FAILED(hi)) { printf("hr=0x%08X init_devices\n", unsigned(hi)); return 1; }
    }
    ScratchImage out;
    HRESULT hr = E_FAIL;
    size_t calls = 0, last = 0; bool monotone = true;
    const size_t abortAt = getenv("TEXCONV_MINI_ABORT_AT") ? strtoull(getenv("TEXCONV_MINI_ABORT_AT"), nullptr, 10) : 0;
    auto cb = [&](size_t done, size_t total) { ++calls; monotone = monotone && done >= last && done <= total; last = done; return !(abortAt && calls >= abortAt); };
    if (op == "compress")
        hr = (items == 1) ? Compress(imgs[0], DXGI_FORMAT(arg), TEX_COMPRESS_FLAGS(flags), TEX_THRESHOLD_DEFAULT, out)
                          : Compress(imgs.data(), items, md, DXGI_FORMAT(arg), TEX_COMPRESS_FLAGS(flags), TEX_THRESHOLD_DEFAULT, out);
    else if (op == "compress_cb")
    {
        CompressOptions o{ TEX_COMPRESS_FLAGS(flags), TEX_THRESHOLD_DEFAULT, TEX_ALPHA_WEIGHT_DEFAULT };
        hr = CompressEx(imgs[0], DXGI_FORMAT(arg), o, out, cb);
        printf("callbacks=%zu last=%zu\n", calls, last);
        if (SUCCEEDED(hr) && (calls < 2 || !monotone || last != h)) hr = E_FAIL;
    }
    else if (op == "convert")
        hr = (items == 1) ? Convert(imgs[0], DXGI_FORMAT(arg), TEX_FILTER_FLAGS(flags), TEX_THRESHOLD_DEFAULT, out)
                          : Convert(imgs.data(), items, md, DXGI_FORMAT(arg), TEX_FILTER_FLAGS(flags), TEX_THRESHOLD_DEFAULT, out);
    else if (op == "mips")
        hr = (items == 1) ? GenerateMipMaps(imgs[0], TEX_FILTER_FLAGS(arg), 0, out)
                          : GenerateMipMaps(imgs.data(), items, md, TEX_FILTER_FLAGS(arg), 0, out);
    else if (op == "decompress")
        hr = Decompress(imgs[0], DXGI_FORMAT(arg), out);
    printf("hr=0x%08X images=%zu bytes=%zu\n", unsigned(hr), out.GetImageCount(), out.GetPixelsSize());
    if (FAILED(hr)) return 1;
    // an output path ending in .dds is written as a DDS file (texconv.cpp:3877 SaveToDDSFile) and read back as a check
    const std::string outPath = argv[3];
    if (outPath.size() > 4 && outPath.compare(outPath.size() - 4, 4, ".dds") == 0)
    {
        hr = SaveToDDSFile(out.GetImages(), out.GetImageCount(), out.GetMetadata(), DDS_FLAGS_NONE, outPath.c_str());
        TexMetadata back{}; ScratchImage re;
        if (SUCCEEDED(hr)) hr = LoadFromDDSFile(outPath.c_str(), DDS_FLAGS_NONE, &back, re);
        if (SUCCEEDED(hr) && (re.GetPixelsSize() != out.GetPixelsSize() || memcmp(re.GetPixels(), out.GetPixels(), out.GetPixelsSize()) != 0)) hr = E_FAIL;
        printf("dds hr=0x%08X\n", unsigned(hr));
        return SUCCEEDED(hr) ? 0 : 5;
    }
    return dump(argv[3], out.GetPixels(), out.GetPixelsSize()) ? 0 : 4;
}
