"""One-off: /root/reference/samples/jfk.mp3 -> tests/golden/jfk_16k.wav (16 kHz mono int16), the input `examples/main -f samples/jfk.wav`
of the reference README expects (README.md:28; read_wav, examples/common.cpp:642-748, takes 16 kHz WAV only and the repo ships the MP3).

No MP3 decoder is written here.  The image has no ffmpeg binary, but opencv-python-headless bundles libavformat / libavcodec
(FFmpeg 8) with the mp3 demuxer + decoder; this script drives them through ctypes (public API calls plus the leading, long-stable
fields of AVFormatContext / AVStream / AVPacket / AVFrame, each sanity-checked), downmixes to mono, resamples to 16 kHz with
scipy.signal.resample_poly and quantises to int16 exactly as a WAV writer does.  It ran once in the build container; the WAV it
wrote is the committed fixture, so nothing at test time needs opencv, scipy or /root/reference."""
import ctypes as C
import glob
import os
import sys
import wave

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/samples/jfk.mp3"
DST = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "jfk_16k.wav")


def load_ffmpeg():
    import cv2  # noqa: F401  (pulls the bundled shared objects and their private dependencies into the process)
    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    libs = {}
    for n in ("libavutil", "libswresample", "libavcodec", "libavformat"):
        libs[n] = C.CDLL(glob.glob(os.path.join(d, n + "-*"))[0], mode=C.RTLD_GLOBAL)
    return libs["libavutil"], libs["libavcodec"], libs["libavformat"]


def decode(path):
    avu, avc, avf = load_ffmpeg()
    vp = C.c_void_p
    avf.avformat_open_input.argtypes = [C.POINTER(vp), C.c_char_p, vp, vp]
    avf.avformat_find_stream_info.argtypes = [vp, vp]
    avf.av_find_best_stream.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.c_int]
    avf.av_read_frame.argtypes = [vp, vp]
    avf.avformat_close_input.argtypes = [C.POINTER(vp)]
    avc.avcodec_alloc_context3.restype = vp
    avc.avcodec_alloc_context3.argtypes = [vp]
    avc.avcodec_parameters_to_context.argtypes = [vp, vp]
    avc.avcodec_open2.argtypes = [vp, vp, vp]
    avc.av_packet_alloc.restype = vp
    avc.av_packet_unref.argtypes = [vp]
    avc.avcodec_send_packet.argtypes = [vp, vp]
    avc.avcodec_receive_frame.argtypes = [vp, vp]
    avu.av_frame_alloc.restype = vp
    avu.av_opt_get_int.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]

    fmt = vp()
    assert avf.avformat_open_input(C.byref(fmt), path.encode(), None, None) == 0, "avformat_open_input failed"
    assert avf.avformat_find_stream_info(fmt, None) >= 0
    dec = vp()
    idx = avf.av_find_best_stream(fmt, 1, -1, -1, C.byref(dec), 0)          # AVMEDIA_TYPE_AUDIO
    assert idx >= 0 and dec.value, "no audio stream / decoder"
    nb_streams = C.c_uint.from_address(fmt.value + 44).value                   # AVFormatContext.nb_streams
    assert 1 <= nb_streams <= 8 and idx < nb_streams, nb_streams
    streams = vp.from_address(fmt.value + 48).value                            # AVFormatContext.streams
    st = vp.from_address(streams + 8 * idx).value
    assert C.c_int.from_address(st + 8).value == idx, "AVStream.index mismatch: struct layout differs"
    par = vp.from_address(st + 16).value                                       # AVStream.codecpar
    cctx = avc.avcodec_alloc_context3(dec)
    assert avc.avcodec_parameters_to_context(cctx, par) >= 0
    assert avc.avcodec_open2(cctx, dec, None) == 0
    ar = C.c_int64()
    assert avu.av_opt_get_int(cctx, b"ar", 0, C.byref(ar)) >= 0 and 8000 <= ar.value <= 192000
    pkt, frame = avc.av_packet_alloc(), avu.av_frame_alloc()
    chunks = []

    def drain():
        while avc.avcodec_receive_frame(cctx, frame) == 0:
            n = C.c_int.from_address(frame + 112).value                        # AVFrame.nb_samples
            f = C.c_int.from_address(frame + 116).value                        # AVFrame.format (enum AVSampleFormat)
            assert 0 < n <= 8192, n
            ext = vp.from_address(frame + 96).value                            # AVFrame.extended_data
            planes = []
            for ch in range(2):
                p = vp.from_address(ext + 8 * ch).value
                if not p:
                    break
                if f == 8:      # fltp
                    planes.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (n,)).astype(np.float64))
                elif f == 6:    # s16p
                    planes.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int16)), (n,)).astype(np.float64) / 32768.0)
                else:
                    raise SystemExit(f"unexpected sample format {f} (expected planar float / planar s16 from the mp3 decoder)")
            chunks.append(np.mean(planes, axis=0))

    while avf.av_read_frame(fmt, pkt) >= 0:
        if C.c_int.from_address(pkt + 36).value == idx:                        # AVPacket.stream_index
            if avc.avcodec_send_packet(cctx, pkt) == 0:
                drain()
        avc.av_packet_unref(pkt)
    avc.avcodec_send_packet(cctx, None)
    drain()
    avf.avformat_close_input(C.byref(fmt))
    return np.concatenate(chunks), int(ar.value)


def main():
    pcm, rate = decode(SRC)
    print(f"decoded {pcm.size} samples at {rate} Hz ({pcm.size / rate:.2f} s), peak {np.abs(pcm).max():.3f}")
    if rate != 16000:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(16000, rate)
        pcm = resample_poly(pcm, 16000 // g, rate // g)
    i16 = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)
    with wave.open(DST, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(16000)
        wf.writeframes(i16.tobytes())
    print(f"wrote {DST}: {i16.size} samples ({i16.size / 16000:.2f} s)")


if __name__ == "__main__":
    main()
