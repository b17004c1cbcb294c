// q2w-main -- the reference CLI (examples/main/main.cpp) on the B200 library, through the public API only.
//
//   q2w-main -m model.bin -f audio.wav [-n iters] [-ot offset_ms] [-d duration_ms] [-dev gpu | -dev -1 (all GPUs)] [--long] [-np]
//
// Mirrors /root/reference/examples/main/main.cpp:353-594 for the part of it that reaches the encoder: read a 16-bit
// 16 kHz mono/stereo WAV (read_wav, examples/common.cpp:642-748: int16 / 32768, stereo averaged), init from file,
// N x { whisper_full ; whisper_print_emb_enc } (the reference hard-codes N = 100 at :574-580; here -n, default 1),
// wall seconds to stderr, whisper_print_timings.  --long adds what the fork dropped: audio longer than 30 s is cut
// into 30 s windows and pushed through whisper_encode_batch (per-window mel normalisation).
// The WAV reader is a small RIFF/WAVE PCM parser (the reference vendors dr_wav.h for this).
#include "qwen2-whisper.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static bool read_wav(const std::string& fname, std::vector<float>& pcmf32) {
    std::vector<uint8_t> d;
    FILE* f = fname == "-" ? stdin : fopen(fname.c_str(), "rb");
    if (!f) { fprintf(stderr, "error: failed to open '%s' as WAV file\n", fname.c_str()); return false; }
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
    if (f != stdin) fclose(f);
    auto u16 = [&](size_t o) { return static_cast<uint32_t>(d[o]) | (static_cast<uint32_t>(d[o + 1]) << 8); };
    auto u32 = [&](size_t o) { return u16(o) | (u16(o + 2) << 16); };
    if (d.size() < 44 || memcmp(d.data(), "RIFF", 4) || memcmp(d.data() + 8, "WAVE", 4)) {
        fprintf(stderr, "error: failed to open '%s' as WAV file\n", fname.c_str());
        return false;
    }
    uint32_t channels = 0, rate = 0, bits = 0, fmt = 0;
    size_t data_off = 0, data_len = 0;
    for (size_t o = 12; o + 8 <= d.size();) {
        const uint32_t len = u32(o + 4);
        if (!memcmp(d.data() + o, "fmt ", 4) && o + 24 <= d.size()) {
            fmt = u16(o + 8); channels = u16(o + 10); rate = u32(o + 12); bits = u16(o + 22);
        } else if (!memcmp(d.data() + o, "data", 4)) {
            data_off = o + 8;
            data_len = std::min<size_t>(len, d.size() - data_off);
            break;
        }
        o += 8 + len + (len & 1);
    }
    if (!data_off || (fmt != 1 && fmt != 0xFFFE)) { fprintf(stderr, "error: '%s' is not a PCM WAV file\n", fname.c_str()); return false; }
    if (channels != 1 && channels != 2) { fprintf(stderr, "read_wav: WAV file '%s' must be mono or stereo\n", fname.c_str()); return false; }
    if (rate != WHISPER_SAMPLE_RATE) { fprintf(stderr, "read_wav: WAV file '%s' must be %i kHz\n", fname.c_str(), WHISPER_SAMPLE_RATE / 1000); return false; }
    if (bits != 16) { fprintf(stderr, "read_wav: WAV file '%s' must be 16-bit\n", fname.c_str()); return false; }
    const size_t frames = data_len / (2 * channels);
    const int16_t* s = reinterpret_cast<const int16_t*>(d.data() + data_off);
    pcmf32.resize(frames);
    if (channels == 1) for (size_t i = 0; i < frames; ++i) pcmf32[i] = float(s[i]) / 32768.0f;
    else for (size_t i = 0; i < frames; ++i) pcmf32[i] = float(s[2 * i] + s[2 * i + 1]) / 65536.0f;
    return true;
}

int main(int argc, char** argv) {
    std::string model = "models/ggml-model-f32.bin";   // the reference's default (examples/main/main.cpp:77)
    std::vector<std::string> files;
    int iters = 1, offset_ms = 0, duration_ms = 0, dev = 0;
    bool no_prints = false, long_mode = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-m" || a == "--model") model = next();
        else if (a == "-f" || a == "--file") files.push_back(next());
        else if (a == "-n" || a == "--iters") iters = atoi(next());
        else if (a == "-ot" || a == "--offset-t") offset_ms = atoi(next());
        else if (a == "-d" || a == "--duration") duration_ms = atoi(next());
        else if (a == "-dev" || a == "--device") dev = atoi(next());
        else if (a == "-t" || a == "--threads") next();      // accepted, ignored (GPU path)
        else if (a == "-np" || a == "--no-prints") no_prints = true;
        else if (a == "--long") long_mode = true;
        else if (a == "-h" || a == "--help") { fprintf(stderr, "usage: %s -m model.bin -f audio.wav [-n iters] [-ot ms] [-d ms] [-dev gpu] [--long] [-np]\n", argv[0]); return 0; }
        else if (a[0] != '-') files.push_back(a);
        else { fprintf(stderr, "error: unknown argument: %s\n", a.c_str()); return 2; }
    }
    if (files.empty()) { fprintf(stderr, "error: no input files specified\n"); return 2; }
    if (no_prints) whisper_log_set([](ggml_log_level, const char*, void*) {}, nullptr);

    whisper_context_params cparams = whisper_context_default_params();
    cparams.gpu_device = dev;
    whisper_context* ctx = whisper_init_from_file_with_params(model.c_str(), cparams);
    if (!ctx) { fprintf(stderr, "error: failed to initialize whisper context\n"); return 3; }
    if (!no_prints) fprintf(stderr, "system_info: %s\n", whisper_print_system_info());

    for (const std::string& fname : files) {
        std::vector<float> pcm;
        if (!read_wav(fname, pcm)) { fprintf(stderr, "error: failed to read WAV file '%s'\n", fname.c_str()); continue; }
        if (!no_prints) fprintf(stderr, "%s: processing '%s' (%zu samples, %.1f sec)\n", argv[0], fname.c_str(), pcm.size(), pcm.size() / 16000.0);
        whisper_full_params wparams = whisper_full_default_params();
        wparams.offset_ms = offset_ms;
        wparams.duration_ms = duration_ms;
        const auto t0 = std::chrono::system_clock::now();
        if (long_mode) {
            const size_t win = static_cast<size_t>(WHISPER_SAMPLE_RATE) * WHISPER_CHUNK_SIZE;
            const int nw = static_cast<int>((pcm.size() + win - 1) / win);
            std::vector<int32_t> ns(nw);
            std::vector<float> padded(static_cast<size_t>(nw) * win, 0.0f);
            memcpy(padded.data(), pcm.data(), pcm.size() * sizeof(float));
            for (int w = 0; w < nw; ++w) ns[w] = static_cast<int32_t>(std::min(win, pcm.size() - static_cast<size_t>(w) * win));
            for (int it = 0; it < iters; ++it) {
                if (whisper_encode_batch(ctx, padded.data(), win, ns.data(), nw, nullptr) != 0) { fprintf(stderr, "%s: failed to process audio\n", argv[0]); return 10; }
                whisper_print_emb_enc(ctx);
            }
            int n_win = 0, n_out = 0, n_state = 0;
            whisper_embd_dims(ctx, &n_win, &n_out, &n_state);
            if (!no_prints) fprintf(stderr, "%s: %d windows x [%d, %d] embeddings\n", argv[0], n_win, n_out, n_state);
        } else {
            for (int it = 0; it < iters; ++it) {
                if (whisper_full(ctx, wparams, pcm.data(), static_cast<int>(pcm.size())) != 0) { fprintf(stderr, "%s: failed to process audio\n", argv[0]); return 10; }
                whisper_print_emb_enc(ctx);
            }
        }
        const std::chrono::duration<double> diff = std::chrono::system_clock::now() - t0;
        fprintf(stderr, "%f\n", diff.count());
    }
    if (!no_prints) whisper_print_timings(ctx);
    whisper_free(ctx);
    return 0;
}
