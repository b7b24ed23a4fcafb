/* Host-side input path (include/sedk_io.h): PCM16 WAV decode with the reference's read_audio semantics
 * (desed_task/dataio/datasets.py:14-74) and pre-decoded int16 shards.  C11 + pthreads, no CUDA. */
#define _GNU_SOURCE
#include "../../include/sedk_io.h"

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static __thread char g_err[512];

static void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* sedkio_last_error(void) { return g_err; }

static uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* walk the RIFF chunks of an open file */
static int probe_fd(int fd, const char* path, sedkio_wav_info* info) {
    unsigned char h[12];
    if (pread(fd, h, 12, 0) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
        set_err("%s: not a RIFF/WAVE file", path);
        return SEDKIO_ERR_FORMAT;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        set_err("%s: fstat failed: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    int64_t pos = 12;
    int have_fmt = 0;
    memset(info, 0, sizeof(*info));
    while (pos + 8 <= (int64_t)st.st_size) {
        unsigned char ch[8];
        if (pread(fd, ch, 8, pos) != 8) break;
        const uint32_t sz = rd32(ch + 4);
        if (memcmp(ch, "fmt ", 4) == 0) {
            unsigned char f[40];
            const uint32_t want = sz < 40 ? sz : 40;
            if (sz < 16 || pread(fd, f, want, pos + 8) != (ssize_t)want) {
                set_err("%s: truncated fmt chunk", path);
                return SEDKIO_ERR_FORMAT;
            }
            uint16_t tag = rd16(f);
            if (tag == 0xFFFE && sz >= 26) tag = rd16(f + 24);        /* WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the tag */
            info->channels = rd16(f + 2);
            info->sample_rate = (int32_t)rd32(f + 4);
            info->bits_per_sample = rd16(f + 14);
            if (tag != 1 || info->bits_per_sample != 16 || info->channels < 1) {
                set_err("%s: only 16-bit integer PCM is supported (format tag %u, %d bits, %d channels)", path, (unsigned)tag,
                        info->bits_per_sample, info->channels);
                return SEDKIO_ERR_FORMAT;
            }
            have_fmt = 1;
        } else if (memcmp(ch, "data", 4) == 0) {
            if (!have_fmt) {
                set_err("%s: data chunk before fmt chunk", path);
                return SEDKIO_ERR_FORMAT;
            }
            int64_t bytes = sz;
            if (pos + 8 + bytes > (int64_t)st.st_size) bytes = (int64_t)st.st_size - pos - 8;      /* streamed / truncated writers */
            info->data_offset = pos + 8;
            info->frames = bytes / (2 * (int64_t)info->channels);
            return SEDKIO_OK;
        }
        pos += 8 + (int64_t)sz + (sz & 1);
    }
    set_err("%s: no data chunk", path);
    return SEDKIO_ERR_FORMAT;
}

int sedkio_wav_probe(const char* path, sedkio_wav_info* info) {
    if (!path || !info) {
        set_err("sedkio_wav_probe: null argument");
        return SEDKIO_ERR_ARG;
    }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_err("%s: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    const int rc = probe_fd(fd, path, info);
    close(fd);
    return rc;
}

/* pad_audio (datasets.py:24-47) on one decoded clip: frames available `have` starting at the cut onset */
static int64_t cut_onset(int64_t frames, int64_t pad_to, const int64_t* onset, int i) {
    if (frames <= pad_to || onset == NULL) return 0;
    int64_t o = onset[i];
    if (o < 0) o = 0;
    if (o > frames - pad_to) o = frames - pad_to;
    return o;
}

typedef struct {
    const char* const* paths;
    int n;
    int64_t pad_to;
    const int64_t* onset;
    const int32_t* channel;
    int16_t* out_i16;
    float* out_f32;
    sedkio_wav_info* info;
    int32_t* status;
    int next;                      /* work counter */
    int first_bad;
    char first_err[512];
    pthread_mutex_t mu;
} read_job;

static int read_one(read_job* j, int i) {
    const char* path = j->paths[i];
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_err("%s: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    sedkio_wav_info wi;
    int rc = probe_fd(fd, path, &wi);
    if (rc != SEDKIO_OK) {
        close(fd);
        return rc;
    }
    if (j->info) j->info[i] = wi;
    const int C = wi.channels;
    const int ch = j->channel ? j->channel[i] : -1;
    if (ch >= C) {
        close(fd);
        set_err("%s: channel %d requested, file has %d", path, ch, C);
        return SEDKIO_ERR_ARG;
    }
    if (j->out_i16 && ch < 0 && C > 1) {
        close(fd);
        set_err("%s: the mean of %d channels is not an int16 signal - ask for the fp32 output or pick a channel", path, C);
        return SEDKIO_ERR_ARG;
    }
    const int64_t T = j->pad_to;
    const int64_t on = cut_onset(wi.frames, T, j->onset, i);
    const int64_t take = wi.frames - on < T ? wi.frames - on : T;
    int16_t* o16 = j->out_i16 ? j->out_i16 + (int64_t)i * T : NULL;
    float* o32 = j->out_f32 ? j->out_f32 + (int64_t)i * T : NULL;
    const int64_t off = wi.data_offset + on * 2 * C;
    if (C == 1 && o16) {
        /* the common case (DESED: 16 kHz mono): straight into the caller's buffer */
        int64_t got = 0;
        while (got < take * 2) {
            const ssize_t r = pread(fd, (char*)o16 + got, (size_t)(take * 2 - got), off + got);
            if (r <= 0) break;
            got += r;
        }
        if (got != take * 2) {
            close(fd);
            set_err("%s: short read", path);
            return SEDKIO_ERR_IO;
        }
        if (o32)
            for (int64_t t = 0; t < take; t++) o32[t] = (float)o16[t] * (1.0f / 32768.0f);
    } else {
        enum { CHUNK = 16384 };
        int16_t* buf = (int16_t*)malloc((size_t)CHUNK * C * 2);
        if (!buf) {
            close(fd);
            set_err("out of memory");
            return SEDKIO_ERR_IO;
        }
        for (int64_t t0 = 0; t0 < take; t0 += CHUNK) {
            const int64_t nt = take - t0 < CHUNK ? take - t0 : CHUNK;
            const int64_t bytes = nt * 2 * C;
            int64_t got = 0;
            while (got < bytes) {
                const ssize_t r = pread(fd, (char*)buf + got, (size_t)(bytes - got), off + t0 * 2 * C + got);
                if (r <= 0) break;
                got += r;
            }
            if (got != bytes) {
                free(buf);
                close(fd);
                set_err("%s: short read", path);
                return SEDKIO_ERR_IO;
            }
            for (int64_t t = 0; t < nt; t++) {
                if (ch >= 0 || C == 1) {
                    const int16_t v = buf[t * C + (ch >= 0 ? ch : 0)];
                    if (o16) o16[t0 + t] = v;
                    if (o32) o32[t0 + t] = (float)v * (1.0f / 32768.0f);
                } else {
                    /* torch.mean over the channel axis of x / 32768: every term and the sum are exact in fp32, then one division */
                    int32_t s = 0;
                    for (int c = 0; c < C; c++) s += buf[t * C + c];
                    o32[t0 + t] = ((float)s * (1.0f / 32768.0f)) / (float)C;
                }
            }
        }
        free(buf);
    }
    close(fd);
    if (o16 && take < T) memset(o16 + take, 0, (size_t)(T - take) * 2);
    if (o32 && take < T) memset(o32 + take, 0, (size_t)(T - take) * 4);
    return SEDKIO_OK;
}

static void* read_worker(void* arg) {
    read_job* j = (read_job*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        const int rc = read_one(j, i);
        if (j->status) j->status[i] = rc;
        if (rc != SEDKIO_OK) {
            pthread_mutex_lock(&j->mu);
            if (j->first_bad < 0 || i < j->first_bad) {
                j->first_bad = i;
                snprintf(j->first_err, sizeof(j->first_err), "%s", g_err);
            }
            pthread_mutex_unlock(&j->mu);
        }
    }
    return NULL;
}

static int pick_threads(int n_threads, int n) {
    if (n_threads <= 0) n_threads = n < 16 ? n : 16;
    if (n_threads > n) n_threads = n;
    if (n_threads > 64) n_threads = 64;
    return n_threads < 1 ? 1 : n_threads;
}

int sedkio_read_audio_batch(const char* const* paths, int n, int64_t pad_to, const int64_t* onset, const int32_t* channel,
                            int16_t* out_i16, float* out_f32, sedkio_wav_info* info, int32_t* status, int n_threads) {
    if (!paths || n < 0 || pad_to <= 0 || (!out_i16 && !out_f32)) {
        set_err("sedkio_read_audio_batch: bad arguments");
        return SEDKIO_ERR_ARG;
    }
    if (n == 0) return SEDKIO_OK;
    read_job j = {paths, n, pad_to, onset, channel, out_i16, out_f32, info, status, 0, -1, {0}, PTHREAD_MUTEX_INITIALIZER};
    const int nt = pick_threads(n_threads, n);
    pthread_t th[64];
    int started = 0;
    for (int t = 1; t < nt; t++)
        if (pthread_create(&th[started], NULL, read_worker, &j) == 0) started++;
    read_worker(&j);                                     /* the calling thread works too */
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    if (j.first_bad >= 0) {
        set_err("%s", j.first_err);
        return status ? status[j.first_bad] : SEDKIO_ERR_IO;
    }
    return SEDKIO_OK;
}

/* ------------------------------------------------------------------------------------------------------------ shards */
static const char MAGIC[8] = {'S', 'E', 'D', 'K', 'P', 'C', 'M', '1'};

struct sedkio_shard {
    const unsigned char* base;
    size_t bytes;
    int64_t n;
    int32_t sample_rate;
    const uint64_t* index;         /* {offset in samples, length} pairs */
    const int16_t* data;
};

int sedkio_shard_write(const char* path, const int16_t* pcm, int64_t stride, const int64_t* lengths, int n, int32_t sample_rate) {
    if (!path || !pcm || !lengths || n < 0 || stride < 0) {
        set_err("sedkio_shard_write: bad arguments");
        return SEDKIO_ERR_ARG;
    }
    FILE* f = fopen(path, "wb");
    if (!f) {
        set_err("%s: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    const uint32_t version = 1, sr = (uint32_t)sample_rate;
    const uint64_t nn = (uint64_t)n;
    int ok = fwrite(MAGIC, 1, 8, f) == 8 && fwrite(&version, 4, 1, f) == 1 && fwrite(&sr, 4, 1, f) == 1 && fwrite(&nn, 8, 1, f) == 1;
    uint64_t off = 0;
    for (int i = 0; ok && i < n; i++) {
        if (lengths[i] < 0 || lengths[i] > stride) {
            fclose(f);
            set_err("sedkio_shard_write: clip %d length %lld outside [0, stride]", i, (long long)lengths[i]);
            return SEDKIO_ERR_ARG;
        }
        const uint64_t e[2] = {off, (uint64_t)lengths[i]};
        ok = fwrite(e, 8, 2, f) == 2;
        off += (uint64_t)lengths[i];
    }
    long pos = ok ? ftell(f) : 0;
    static const char zeros[64] = {0};
    if (ok && pos % 64) ok = fwrite(zeros, 1, (size_t)(64 - pos % 64), f) == (size_t)(64 - pos % 64);
    for (int i = 0; ok && i < n; i++)
        if (lengths[i] > 0) ok = fwrite(pcm + (int64_t)i * stride, 2, (size_t)lengths[i], f) == (size_t)lengths[i];
    if (fclose(f) != 0) ok = 0;
    if (!ok) {
        set_err("%s: write failed: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    return SEDKIO_OK;
}

int sedkio_shard_open(const char* path, sedkio_shard** out) {
    if (!path || !out) {
        set_err("sedkio_shard_open: null argument");
        return SEDKIO_ERR_ARG;
    }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_err("%s: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) {
        close(fd);
        set_err("%s: too short for a shard", path);
        return SEDKIO_ERR_FORMAT;
    }
    void* m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
        set_err("%s: mmap failed: %s", path, strerror(errno));
        return SEDKIO_ERR_IO;
    }
    const unsigned char* b = (const unsigned char*)m;
    uint64_t n;
    memcpy(&n, b + 16, 8);
    size_t hdr = 24 + (size_t)n * 16;
    if (hdr % 64) hdr += 64 - hdr % 64;
    int bad = memcmp(b, MAGIC, 8) != 0 || rd32(b + 8) != 1 || n > ((uint64_t)st.st_size - 24) / 16 || hdr > (size_t)st.st_size;
    const uint64_t* index = (const uint64_t*)(b + 24);
    const uint64_t avail = bad ? 0 : ((uint64_t)st.st_size - hdr) / 2;
    for (uint64_t i = 0; !bad && i < n; i++)
        if (index[2 * i] > avail || index[2 * i + 1] > avail - index[2 * i]) bad = 1;
    if (bad) {
        munmap(m, (size_t)st.st_size);
        set_err("%s: not a SEDKPCM1 shard (or truncated)", path);
        return SEDKIO_ERR_FORMAT;
    }
    sedkio_shard* s = (sedkio_shard*)calloc(1, sizeof(*s));
    if (!s) {
        munmap(m, (size_t)st.st_size);
        set_err("out of memory");
        return SEDKIO_ERR_IO;
    }
    s->base = b;
    s->bytes = (size_t)st.st_size;
    s->n = (int64_t)n;
    s->sample_rate = (int32_t)rd32(b + 12);
    s->index = index;
    s->data = (const int16_t*)(b + hdr);
    *out = s;
    return SEDKIO_OK;
}

void sedkio_shard_close(sedkio_shard* s) {
    if (!s) return;
    munmap((void*)s->base, s->bytes);
    free(s);
}
int64_t sedkio_shard_clips(const sedkio_shard* s) { return s ? s->n : 0; }
int32_t sedkio_shard_sample_rate(const sedkio_shard* s) { return s ? s->sample_rate : 0; }
int64_t sedkio_shard_length(const sedkio_shard* s, int64_t clip) {
    return (s && clip >= 0 && clip < s->n) ? (int64_t)s->index[2 * clip + 1] : -1;
}

typedef struct {
    const sedkio_shard* s;
    const int64_t* idx;
    int n;
    int64_t pad_to;
    const int64_t* onset;
    int16_t* out;
    int next;
    pthread_mutex_t mu;
} gather_job;

static void* gather_worker(void* arg) {
    gather_job* j = (gather_job*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        const int64_t c = j->idx[i];
        const int64_t len = (int64_t)j->s->index[2 * c + 1];
        const int64_t on = cut_onset(len, j->pad_to, j->onset, i);
        const int64_t take = len - on < j->pad_to ? len - on : j->pad_to;
        int16_t* o = j->out + (int64_t)i * j->pad_to;
        memcpy(o, j->s->data + j->s->index[2 * c] + on, (size_t)take * 2);
        if (take < j->pad_to) memset(o + take, 0, (size_t)(j->pad_to - take) * 2);
    }
    return NULL;
}

int sedkio_shard_gather(const sedkio_shard* s, const int64_t* idx, int n, int64_t pad_to, const int64_t* onset, int16_t* out,
                        int n_threads) {
    if (!s || !idx || !out || n < 0 || pad_to <= 0) {
        set_err("sedkio_shard_gather: bad arguments");
        return SEDKIO_ERR_ARG;
    }
    for (int i = 0; i < n; i++)
        if (idx[i] < 0 || idx[i] >= s->n) {
            set_err("sedkio_shard_gather: clip index %lld outside [0, %lld)", (long long)idx[i], (long long)s->n);
            return SEDKIO_ERR_ARG;
        }
    if (n == 0) return SEDKIO_OK;
    gather_job j = {s, idx, n, pad_to, onset, out, 0, PTHREAD_MUTEX_INITIALIZER};
    const int nt = pick_threads(n_threads, n);
    pthread_t th[64];
    int started = 0;
    for (int t = 1; t < nt; t++)
        if (pthread_create(&th[started], NULL, gather_worker, &j) == 0) started++;
    gather_worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return SEDKIO_OK;
}
