"""Regenerates tests/golden/tone_stereo_22k.m4a (AAC-LC in an MP4 container) and tests/golden/tone_stereo_22k.flac with the
FFmpeg 8 libavformat / libavcodec that ship in this image (no ffmpeg binary here, and the encoders are only needed to make
the fixtures; the product only DECODES, audio_processor_b200/avdecode.py).

    python tests/golden/make_compressed_fixtures.py

The few struct fields an encoder needs that have no AVOption (AVCodecContext.sample_fmt / frame_size, AVFrame.sample_rate /
ch_layout) are located at run time: AVOptions publish the offsets of their neighbours, and a decoded frame of a WAV with
an unusual rate shows where the decoder put sample_rate and the channel layout."""
import ctypes as C
import os
import sys
import tempfile
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from audio_processor_b200 import avdecode as ad  # noqa: E402

RATE, SECS = 22050, 1.5


def signal():
    t = np.arange(int(RATE * SECS)) / RATE
    l = 0.30 * np.sin(2 * np.pi * 440 * t) + 0.10 * np.sin(2 * np.pi * 1330 * t)
    r = 0.25 * np.sin(2 * np.pi * 660 * t) * (t > 0.4)
    return np.stack([l, r], axis=1).astype(np.float32)


class AVOption(C.Structure):
    _fields_ = [("name", C.c_char_p), ("help", C.c_char_p), ("offset", C.c_int), ("type", C.c_int)]


def option_offsets(avu, obj):
    avu.av_opt_next.argtypes = [C.c_void_p, C.c_void_p]
    avu.av_opt_next.restype = C.POINTER(AVOption)
    out, prev = {}, None
    while True:
        o = avu.av_opt_next(obj, prev)
        if not o:
            return out
        out.setdefault(o.contents.name.decode(), o.contents.offset)
        prev = o


def frame_offsets(avu, avc, avf):
    """(offset of AVFrame.sample_rate, offset of AVFrame.ch_layout): decode a stereo WAV at 12 345 Hz and look for them"""
    d = tempfile.mkdtemp()
    p = os.path.join(d, "probe.wav")
    w = wave.open(p, "wb"); w.setnchannels(2); w.setsampwidth(2); w.setframerate(12345); w.writeframes(np.zeros((4096, 2), np.int16).tobytes()); w.close()
    fmt = C.c_void_p(None)
    assert avf.avformat_open_input(C.byref(fmt), p.encode(), None, None) == 0
    avf.avformat_find_stream_info(fmt, None)
    dec = C.c_void_p(None)
    idx = avf.av_find_best_stream(fmt, 1, -1, -1, C.byref(dec), 0)
    streams = C.c_void_p.from_address(fmt.value + 48).value
    st = C.c_void_p.from_address(streams + 8 * idx).value
    par = C.c_void_p.from_address(st + 16).value
    cctx = C.c_void_p(avc.avcodec_alloc_context3(dec))
    avc.avcodec_parameters_to_context(cctx, par); avc.avcodec_open2(cctx, dec, None)
    pkt = C.c_void_p(avc.av_packet_alloc()); frame = C.c_void_p(avu.av_frame_alloc())
    assert avf.av_read_frame(fmt, pkt) >= 0 and avc.avcodec_send_packet(cctx, pkt) >= 0 and avc.avcodec_receive_frame(cctx, frame) >= 0
    raw = C.string_at(frame.value, 1024)                       # sizeof(AVFrame) is > 1 KB in FFmpeg 8 (the struct ends with ch_layout, duration)
    rate_off = [o for o in range(120, 1000, 4) if int.from_bytes(raw[o:o + 4], "little") == 12345]
    # AVChannelLayout {order, nb_channels = 2, mask, opaque = NULL}: a WAV without a channel mask decodes with order UNSPEC (0),
    # mask 0; it sits behind the buffer pointers / flags (> 300) and is followed by the frame duration (= nb_samples here)
    def is_layout(o):
        order, nbc = int.from_bytes(raw[o:o + 4], "little"), int.from_bytes(raw[o + 4:o + 8], "little")
        mask, opq = int.from_bytes(raw[o + 8:o + 16], "little"), int.from_bytes(raw[o + 16:o + 24], "little")
        return order in (0, 1) and nbc == 2 and mask in (0, 3) and opq == 0
    lay_off = [o for o in range(304, 520, 8) if is_layout(o)]
    assert len(rate_off) == 1 and len(lay_off) == 1, (rate_off, lay_off)
    avc.avcodec_free_context(C.byref(cctx)); avf.avformat_close_input(C.byref(fmt))
    return rate_off[0], lay_off[0]


def encode(path, codec_name, sample_fmt, pcm_f32):
    avu, avc, avf = ad._load()
    P, PP = C.c_void_p, C.POINTER(C.c_void_p)
    avf.avformat_alloc_output_context2.argtypes = [PP, P, C.c_char_p, C.c_char_p]
    avf.avformat_new_stream.argtypes = [P, P]; avf.avformat_new_stream.restype = P
    avf.avio_open.argtypes = [PP, C.c_char_p, C.c_int]
    avf.avformat_write_header.argtypes = [P, P]; avf.av_interleaved_write_frame.argtypes = [P, P]; avf.av_write_trailer.argtypes = [P]
    avf.avio_closep.argtypes = [PP]; avf.avformat_free_context.argtypes = [P]
    avc.avcodec_find_encoder_by_name.argtypes = [C.c_char_p]; avc.avcodec_find_encoder_by_name.restype = P
    avc.avcodec_parameters_from_context.argtypes = [P, P]
    avc.avcodec_send_frame.argtypes = [P, P]; avc.avcodec_receive_packet.argtypes = [P, P]
    avc.av_packet_rescale_ts.argtypes = [P, C.c_int64, C.c_int64]          # two AVRational by value = two 8-byte ints
    avu.av_opt_set_int.argtypes = [P, C.c_char_p, C.c_int64, C.c_int]
    avu.av_opt_set_chlayout.argtypes = [P, C.c_char_p, C.POINTER(ad._AVChannelLayout), C.c_int]
    avu.av_channel_layout_default.argtypes = [C.POINTER(ad._AVChannelLayout), C.c_int]
    rate_off, lay_off = frame_offsets(avu, avc, avf)

    oc = C.c_void_p(None)
    assert avf.avformat_alloc_output_context2(C.byref(oc), None, None, path.encode()) >= 0
    enc = avc.avcodec_find_encoder_by_name(codec_name)
    assert enc
    st = avf.avformat_new_stream(oc, None)
    cctx = C.c_void_p(avc.avcodec_alloc_context3(enc))
    offs = option_offsets(avu, cctx)
    assert offs["ch_layout"] - offs["ar"] == 8, offs            # int sample_rate; enum AVSampleFormat sample_fmt; AVChannelLayout ch_layout; int frame_size
    lay = ad._AVChannelLayout(); avu.av_channel_layout_default(C.byref(lay), 2)
    assert avu.av_opt_set_int(cctx, b"ar", RATE, 0) >= 0 and avu.av_opt_set_chlayout(cctx, b"ch_layout", C.byref(lay), 0) >= 0
    avu.av_opt_set_int(cctx, b"b", 96000, 0)
    C.c_int32.from_address(cctx.value + offs["ar"] + 4).value = sample_fmt
    avu.av_opt_set_int(cctx, b"flags", 1 << 22, 0)              # AV_CODEC_FLAG_GLOBAL_HEADER: mp4 wants the AudioSpecificConfig in extradata
    assert avc.avcodec_open2(cctx, enc, None) >= 0
    frame_size = C.c_int32.from_address(cctx.value + offs["ch_layout"] + 24).value
    if not (0 < frame_size <= 8192):
        frame_size = 1024 if codec_name == b"aac" else 4096
    par = C.c_void_p.from_address(st + 16).value
    assert avc.avcodec_parameters_from_context(par, cctx) >= 0
    pb = C.c_void_p.from_address(oc.value + 32)
    assert avf.avio_open(C.byref(pb), path.encode(), 2) >= 0   # AVIO_FLAG_WRITE; oc->pb is the fifth pointer
    assert avf.avformat_write_header(oc, None) >= 0
    tb_st = C.c_int64.from_address(st + 32).value               # AVStream.time_base (AVRational) follows priv_data
    tb_enc = (1 & 0xffffffff) | (RATE << 32)                    # AVRational{1, RATE} passed as one 64-bit register
    pkt = C.c_void_p(avc.av_packet_alloc())
    frame = C.c_void_p(avu.av_frame_alloc())
    n = len(pcm_f32)
    planar = sample_fmt >= 5
    pts = 0

    def drain():
        while avc.avcodec_receive_packet(cctx, pkt) >= 0:
            avc.av_packet_rescale_ts(pkt, tb_enc, tb_st)
            C.c_int32.from_address(pkt.value + 36).value = 0   # stream_index
            assert avf.av_interleaved_write_frame(oc, pkt) >= 0

    for s in range(0, n, frame_size):
        blk = pcm_f32[s:s + frame_size]
        k = len(blk)
        if sample_fmt in (1, 6):
            blk = np.clip(np.rint(blk * 32768.0), -32768, 32767).astype(np.int16)
        bufs = [np.ascontiguousarray(blk[:, c]) for c in range(2)] if planar else [np.ascontiguousarray(blk)]
        avu.av_frame_unref(frame)
        for c, b in enumerate(bufs):
            C.c_void_p.from_address(frame.value + 8 * c).value = b.ctypes.data        # data[c]
            C.c_int32.from_address(frame.value + 64 + 4 * c).value = b.nbytes         # linesize[c]
        C.c_void_p.from_address(frame.value + 96).value = frame.value                 # extended_data = data
        C.c_int32.from_address(frame.value + 112).value = k
        C.c_int32.from_address(frame.value + 116).value = sample_fmt
        C.c_int32.from_address(frame.value + rate_off).value = RATE
        C.memmove(frame.value + lay_off, C.byref(lay), 24)
        C.c_int64.from_address(frame.value + PTS_OFF[0]).value = pts
        pts += k
        assert avc.avcodec_send_frame(cctx, frame) >= 0, "avcodec_send_frame"
        drain()
        del bufs
    avc.avcodec_send_frame(cctx, None)
    drain()
    assert avf.av_write_trailer(oc) >= 0
    avf.avio_closep(C.byref(pb))
    print("wrote", path, os.path.getsize(path), "bytes")


PTS_OFF = [None]


def find_pts_offset():
    """AVFrame.pts: decode the probe WAV's SECOND frame; its pts is the number of samples of the first (time base 1 / rate)"""
    avu, avc, avf = ad._load()
    d = tempfile.mkdtemp(); p = os.path.join(d, "probe.wav")
    w = wave.open(p, "wb"); w.setnchannels(1); w.setsampwidth(2); w.setframerate(12345); w.writeframes(np.zeros(30000, np.int16).tobytes()); w.close()
    fmt = C.c_void_p(None); avf.avformat_open_input(C.byref(fmt), p.encode(), None, None); avf.avformat_find_stream_info(fmt, None)
    dec = C.c_void_p(None); idx = avf.av_find_best_stream(fmt, 1, -1, -1, C.byref(dec), 0)
    streams = C.c_void_p.from_address(fmt.value + 48).value; st = C.c_void_p.from_address(streams + 8 * idx).value
    cctx = C.c_void_p(avc.avcodec_alloc_context3(dec)); avc.avcodec_parameters_to_context(cctx, C.c_void_p.from_address(st + 16).value); avc.avcodec_open2(cctx, dec, None)
    pkt = C.c_void_p(avc.av_packet_alloc()); frame = C.c_void_p(avu.av_frame_alloc())
    got = []
    while len(got) < 2 and avf.av_read_frame(fmt, pkt) >= 0:
        avc.avcodec_send_packet(cctx, pkt)
        while avc.avcodec_receive_frame(cctx, frame) >= 0 and len(got) < 2:
            got.append((C.c_int32.from_address(frame.value + 112).value, C.string_at(frame.value, 256)))
        avc.av_packet_unref(pkt)
    n0, raw1 = got[0][0], got[1][1]
    cands = [o for o in range(120, 240, 8) if int.from_bytes(raw1[o:o + 8], "little", signed=True) == n0]
    assert cands, "pts offset not found"
    PTS_OFF[0] = cands[0]


if __name__ == "__main__":
    find_pts_offset()
    x = signal()
    encode(os.path.join(HERE, "tone_stereo_22k.m4a"), b"aac", 8, x)        # AV_SAMPLE_FMT_FLTP
    encode(os.path.join(HERE, "tone_stereo_22k.flac"), b"flac", 1, x)      # AV_SAMPLE_FMT_S16
    for name in ("tone_stereo_22k.m4a", "tone_stereo_22k.flac"):
        pcm, rate = ad.decode_audio(os.path.join(HERE, name))
        print(name, pcm.shape, pcm.dtype, rate, float(np.abs(pcm.astype(np.float64)).max()))
