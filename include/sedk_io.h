/* sedk_io.h - host-side input path of the SED hot loop: PCM16 WAV decoding with the reference's
 * read_audio semantics, and pre-decoded int16 shards.  Plain C ABI, no CUDA: the buffers these calls fill (ideally pinned
 * host memory) are what the engines copy to the device as 16-bit PCM (sedk_logmel_fwd_i16 in sedk.h: x / 32768 in the
 * front end's load path, bit-identical to the normalised fp32 waveform torchaudio.load returns).
 *
 * Replaces, per batch instead of per item in DataLoader workers:
 *   desed_task/dataio/datasets.py:57-74  read_audio  = torchaudio.load -> to_mono (:14-21) -> pad_audio (:24-47) -> float()
 * Built by `python -m desed_task_b200.build` into desed_task_b200/lib/libsedkio.so (gcc, pthreads).
 */
#ifndef SEDK_IO_H
#define SEDK_IO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SEDKIO_API __attribute__((visibility("default")))
#else
#define SEDKIO_API
#endif

#define SEDKIO_OK 0
#define SEDKIO_ERR_IO 1            /* open / read / mmap failed */
#define SEDKIO_ERR_FORMAT 2        /* not RIFF/WAVE, truncated, or not 16-bit integer PCM */
#define SEDKIO_ERR_ARG 3

typedef struct {
    int32_t sample_rate, channels, bits_per_sample;
    int64_t frames;                /* samples per channel */
    int64_t data_offset;           /* byte offset of the first sample in the file */
} sedkio_wav_info;

/* thread-local description of the last failure of the calling thread's own call (batch calls: first failing file) */
SEDKIO_API const char* sedkio_last_error(void);

SEDKIO_API int sedkio_wav_probe(const char* path, sedkio_wav_info* info);

/* read_audio (datasets.py:57-74) for n files at once, on n_threads threads (<= 0: one per file up to 16).
 *   pad_to     target length in samples (pad_audio's target_len; > 0): shorter clips are zero-padded at the end, longer ones
 *              are cut to [onset[i], onset[i] + pad_to)  (onset NULL or test mode: 0; the caller draws random.randint(0,
 *              frames - pad_to) like pad_audio:36 so the reference's RNG stream is kept);
 *   channel    per file: -1 = mean over the channels (to_mono:17), c >= 0 = that channel (to_mono:19-20; caller draws);
 *              NULL = -1 for every file;
 *   out_i16    [n, pad_to] or NULL: exact for one channel / a picked channel; refused (SEDKIO_ERR_ARG) for the mean of >= 2
 *              channels, which is not an integer in general;
 *   out_f32    [n, pad_to] or NULL: the waveform torchaudio.load would give (sample / 32768), channel mean as sum / channels;
 *   info       [n] or NULL: the probe result of every file (frames BEFORE cutting: the caller computes onset_s / offset_s);
 *   status     [n] or NULL: per-file SEDKIO_* code.  Returns SEDKIO_OK only if every file succeeded. */
SEDKIO_API int sedkio_read_audio_batch(const char* const* paths, int n, int64_t pad_to, const int64_t* onset,
                                       const int32_t* channel, int16_t* out_i16, float* out_f32, sedkio_wav_info* info,
                                       int32_t* status, int n_threads);

/* ---- pre-decoded shards: "SEDKPCM1" | u32 version | u32 sample_rate | u64 n_clips | {u64 offset, u64 length}[n] | pad to 64
 *      bytes | int16 samples (mono, little endian).  One file per few thousand clips replaces per-clip open + decode. */
SEDKIO_API int sedkio_shard_write(const char* path, const int16_t* pcm, int64_t stride, const int64_t* lengths, int n,
                                  int32_t sample_rate);
typedef struct sedkio_shard sedkio_shard;
SEDKIO_API int sedkio_shard_open(const char* path, sedkio_shard** out);       /* mmap, read-only */
SEDKIO_API void sedkio_shard_close(sedkio_shard* s);
SEDKIO_API int64_t sedkio_shard_clips(const sedkio_shard* s);
SEDKIO_API int32_t sedkio_shard_sample_rate(const sedkio_shard* s);
SEDKIO_API int64_t sedkio_shard_length(const sedkio_shard* s, int64_t clip);  /* -1: bad index */
/* out [n, pad_to] <- clips idx[0..n) with pad_audio's pad / cut rule (onset as above) */
SEDKIO_API int sedkio_shard_gather(const sedkio_shard* s, const int64_t* idx, int n, int64_t pad_to, const int64_t* onset,
                                   int16_t* out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
