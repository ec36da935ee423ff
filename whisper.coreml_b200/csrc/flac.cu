// Host-side FLAC decoder (no CUDA in this file): the reference reads compressed audio through an `ffmpeg` subprocess
// (whisper/audio.py:45-62); this backend decodes the container itself and hands the PCM to the device resampler
// (resample.cu).  FLAC is lossless: the decoded samples are checked bit for bit against the MD5 the encoder stored in
// STREAMINFO (tests/test_audio_ingest.py).  Entropy decoding is a serial bit stream (Rice codes), so it stays on the host;
// everything after it (channel mix, resampling, log-mel) runs on the GPU.
//
// Format (FLAC format specification, "frame" / "subframe" / "residual"): 'fLaC', metadata blocks (STREAMINFO first), then
// frames = header (sync 0x3FFE, block size / sample rate / channel assignment / sample size codes, UTF-8 coded number,
// CRC-8), one subframe per channel (constant | verbatim | fixed predictor order 0-4 | LPC order 1-32, Rice-coded residual in
// 2^order partitions), zero padding to a byte boundary, CRC-16.  Stereo decorrelation: left/side, side/right, mid/side.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "whisper_b200.h"

namespace b200 {
namespace {

struct BitReader {
    const uint8_t* p; long n; long pos = 0;     // pos in bits
    bool fail = false;
    uint32_t bits(int k) {                       // k <= 32, MSB first
        uint32_t v = 0;
        while (k > 0) {
            const long byte = pos >> 3;
            if (byte >= n) { fail = true; return 0; }
            const int avail = 8 - (int)(pos & 7), take = k < avail ? k : avail;
            v = (v << take) | ((p[byte] >> (avail - take)) & ((1u << take) - 1u));
            pos += take; k -= take;
        }
        return v;
    }
    int32_t sbits(int k) {                       // two's complement, k <= 32
        if (k == 0) return 0;
        const uint32_t v = bits(k);
        return k == 32 ? (int32_t)v : (int32_t)(v << (32 - k)) >> (32 - k);
    }
    uint32_t unary() {                           // zeros before the next 1 bit
        uint32_t z = 0;
        for (;;) {
            const long byte = pos >> 3;
            if (byte >= n) { fail = true; return 0; }
            const int off = (int)(pos & 7);
            const uint32_t rest = (uint32_t)((p[byte] << off) & 0xff);             // remaining bits of this byte, left aligned
            if (rest == 0) { z += 8 - off; pos += 8 - off; continue; }
            const int lead = __builtin_clz(rest) - 24;
            z += lead; pos += lead + 1;
            return z;
        }
    }
    void align() { pos = (pos + 7) & ~7L; }
};

struct StreamInfo { int rate = 0, channels = 0, bps = 0; long total = 0; int min_block = 0, max_block = 0; uint8_t md5[16] = {0}; long first_frame = 0; };

bool parse_header(const uint8_t* d, long n, StreamInfo* si) {
    if (n < 42 || memcmp(d, "fLaC", 4) != 0) return false;
    long o = 4;
    bool have = false;
    for (;;) {
        if (o + 4 > n) return false;
        const bool last = d[o] & 0x80;
        const int type = d[o] & 0x7f;
        const long len = ((long)d[o + 1] << 16) | ((long)d[o + 2] << 8) | d[o + 3];
        o += 4;
        if (o + len > n) return false;
        if (type == 0 && len >= 34) {
            const uint8_t* s = d + o;
            si->min_block = (s[0] << 8) | s[1]; si->max_block = (s[2] << 8) | s[3];
            si->rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            si->channels = ((s[12] >> 1) & 7) + 1;
            si->bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            si->total = ((long)(s[13] & 0xf) << 32) | ((long)s[14] << 24) | ((long)s[15] << 16) | ((long)s[16] << 8) | s[17];
            memcpy(si->md5, s + 18, 16);
            have = true;
        }
        o += len;
        if (last) break;
    }
    si->first_frame = o;
    return have;
}

bool read_residual(BitReader& br, int32_t* out, int blocksize, int order) {
    const int method = (int)br.bits(2);
    if (method > 1) return false;
    const int pbits = method == 0 ? 4 : 5, escape = method == 0 ? 15 : 31;
    const int porder = (int)br.bits(4), parts = 1 << porder;
    if ((blocksize >> porder) << porder != blocksize && porder > 0) return false;
    int idx = order;
    for (int p = 0; p < parts; ++p) {
        const int count = (blocksize >> porder) - (p == 0 ? order : 0);
        if (count < 0) return false;
        const int k = (int)br.bits(pbits);
        if (k == escape) {
            const int raw = (int)br.bits(5);
            for (int i = 0; i < count; ++i) out[idx++] = br.sbits(raw);
        } else {
            for (int i = 0; i < count; ++i) {
                const uint32_t q = br.unary();
                const uint32_t u = (q << k) | (k ? br.bits(k) : 0u);
                out[idx++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
            }
        }
        if (br.fail) return false;
    }
    return idx == blocksize;
}

bool read_subframe(BitReader& br, int32_t* out, int blocksize, int bps) {
    if (br.bits(1) != 0) return false;
    const int type = (int)br.bits(6);
    int wasted = 0;
    if (br.bits(1)) wasted = (int)br.unary() + 1;
    bps -= wasted;
    if (bps < 1 || bps > 33) return false;
    auto sample = [&](int b) -> int64_t {          // up to 33 bits (side channel of 32-bit audio): two reads
        if (b <= 32) return br.sbits(b);
        const int64_t hi = br.sbits(b - 32);
        return (hi << 32) | br.bits(32);
    };
    if (type == 0) {                                // constant
        const int32_t v = (int32_t)sample(bps);
        for (int i = 0; i < blocksize; ++i) out[i] = v;
    } else if (type == 1) {                         // verbatim
        for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)sample(bps);
    } else if (type >= 8 && type <= 12) {           // fixed predictor
        const int order = type - 8;
        if (order > blocksize) return false;
        for (int i = 0; i < order; ++i) out[i] = (int32_t)sample(bps);
        if (!read_residual(br, out, blocksize, order)) return false;
        for (int i = order; i < blocksize; ++i) {
            int64_t pred = 0;
            switch (order) {
            case 1: pred = out[i - 1]; break;
            case 2: pred = 2 * (int64_t)out[i - 1] - out[i - 2]; break;
            case 3: pred = 3 * (int64_t)out[i - 1] - 3 * (int64_t)out[i - 2] + out[i - 3]; break;
            case 4: pred = 4 * (int64_t)out[i - 1] - 6 * (int64_t)out[i - 2] + 4 * (int64_t)out[i - 3] - out[i - 4]; break;
            default: break;
            }
            out[i] = (int32_t)(out[i] + pred);
        }
    } else if (type >= 32) {                        // LPC
        const int order = (type & 31) + 1;
        if (order > blocksize) return false;
        for (int i = 0; i < order; ++i) out[i] = (int32_t)sample(bps);
        const int prec = (int)br.bits(4) + 1;
        if (prec == 16) return false;
        const int shift = br.sbits(5);
        if (shift < 0) return false;
        int32_t coef[32];
        for (int i = 0; i < order; ++i) coef[i] = br.sbits(prec);
        if (!read_residual(br, out, blocksize, order)) return false;
        for (int i = order; i < blocksize; ++i) {
            int64_t sum = 0;
            for (int j = 0; j < order; ++j) sum += (int64_t)coef[j] * out[i - 1 - j];
            out[i] = (int32_t)(out[i] + (sum >> shift));
        }
    } else {
        return false;                               // reserved
    }
    if (wasted)
        for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)((uint32_t)out[i] << wasted);
    return !br.fail;
}

// one frame starting at byte offset `o`; appends `blocksize` samples per channel to out[ch]; returns bytes consumed or -1
long read_frame(const uint8_t* d, long n, long o, const StreamInfo& si, std::vector<std::vector<int32_t>>& out) {
    BitReader br{d + o, n - o};
    if (br.bits(14) != 0x3FFE) return -1;
    br.bits(1);
    br.bits(1);                                     // blocking strategy: only the coded number's meaning changes
    const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4), ch_code = (int)br.bits(4), ss_code = (int)br.bits(3);
    br.bits(1);
    {   // UTF-8 style coded frame / sample number (1-7 bytes)
        const uint32_t b0 = br.bits(8);
        int extra = 0;
        if (b0 & 0x80) { for (int m = 0x40; b0 & m; m >>= 1) ++extra; }
        for (int i = 0; i < extra; ++i) br.bits(8);
    }
    int blocksize;
    if (bs_code == 1) blocksize = 192;
    else if (bs_code >= 2 && bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = (int)br.bits(8) + 1;
    else if (bs_code == 7) blocksize = (int)br.bits(16) + 1;
    else if (bs_code >= 8) blocksize = 256 << (bs_code - 8);
    else return -1;
    if (sr_code == 12) br.bits(8); else if (sr_code == 13 || sr_code == 14) br.bits(16); else if (sr_code == 15) return -1;
    static const int ss_table[8] = {0, 8, 12, 0, 16, 20, 24, 32};
    const int bps = ss_code == 0 ? si.bps : ss_table[ss_code];
    if (bps == 0) return -1;
    br.bits(8);                                     // CRC-8 (the stream is validated by the MD5 of all samples instead)
    const int nch = ch_code < 8 ? ch_code + 1 : 2;
    if (nch != si.channels || ch_code > 10 || br.fail) return -1;
    std::vector<int32_t> buf[8];
    for (int c = 0; c < nch; ++c) {
        buf[c].resize(blocksize);
        const int side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
        if (!read_subframe(br, buf[c].data(), blocksize, bps + side)) return -1;
    }
    if (ch_code == 8) { for (int i = 0; i < blocksize; ++i) buf[1][i] = buf[0][i] - buf[1][i]; }              // left, side
    else if (ch_code == 9) { for (int i = 0; i < blocksize; ++i) buf[0][i] = buf[0][i] + buf[1][i]; }         // side, right
    else if (ch_code == 10) {                                                                                  // mid, side
        for (int i = 0; i < blocksize; ++i) {
            const int32_t side = buf[1][i];
            const int32_t mid = (int32_t)(((uint32_t)buf[0][i] << 1) | (uint32_t)(side & 1));
            buf[0][i] = (mid + side) >> 1; buf[1][i] = (mid - side) >> 1;
        }
    }
    br.align();
    br.bits(16);                                    // CRC-16
    if (br.fail) return -1;
    for (int c = 0; c < nch; ++c) out[c].insert(out[c].end(), buf[c].begin(), buf[c].end());
    return br.pos >> 3;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200FlacInfo(const unsigned char* data, long n_bytes, int* sample_rate, int* channels, int* bits_per_sample, long* total_samples,
                 unsigned char* md5_16) {
    StreamInfo si;
    if (!data || !parse_header(data, n_bytes, &si)) { record_error("b200FlacInfo: not a FLAC stream (no 'fLaC' marker / STREAMINFO)"); return -1; }
    if (sample_rate) *sample_rate = si.rate;
    if (channels) *channels = si.channels;
    if (bits_per_sample) *bits_per_sample = si.bps;
    if (total_samples) *total_samples = si.total;
    if (md5_16) memcpy(md5_16, si.md5, 16);
    return 0;
}

long b200FlacDecode(const unsigned char* data, long n_bytes, int* out_interleaved, long cap_samples_per_channel) {
    StreamInfo si;
    if (!data || !parse_header(data, n_bytes, &si)) { record_error("b200FlacDecode: not a FLAC stream"); return -1; }
    std::vector<std::vector<int32_t>> ch(si.channels);
    for (auto& c : ch) c.reserve(si.total > 0 ? (size_t)si.total : 1 << 20);
    long o = si.first_frame;
    while (o + 2 <= n_bytes) {
        if (!(data[o] == 0xFF && (data[o + 1] & 0xFC) == 0xF8)) { ++o; continue; }       // resynchronise on the frame sync code
        const long used = read_frame(data, n_bytes, o, si, ch);
        if (used < 0) { ++o; continue; }                                                  // a false sync inside other data
        o += used;
    }
    const long got = (long)ch[0].size();
    if (si.total > 0 && got != si.total) { record_error("b200FlacDecode: decoded %ld samples per channel, STREAMINFO says %ld", got, si.total); return -1; }
    if (got > cap_samples_per_channel) { record_error("b200FlacDecode: %ld samples per channel exceed the buffer (%ld)", got, cap_samples_per_channel); return -1; }
    for (long i = 0; i < got; ++i)
        for (int c = 0; c < si.channels; ++c) out_interleaved[i * si.channels + c] = ch[c][i];
    return got;
}

}  // extern "C"
