/* fm_wav.c -- see include/fm_wav.h. */
#include "fm_wav.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct fm_wav {
    FILE *f;
    int is_stdout, keep_tail;
    size_t fill;
    unsigned char cluster[FM_WAV_CLUSTER_BYTES];
};

static void le32(unsigned char *p, uint32_t v)
{
    p[0] = (unsigned char) v; p[1] = (unsigned char) (v >> 8); p[2] = (unsigned char) (v >> 16); p[3] = (unsigned char) (v >> 24);
}
static void le16(unsigned char *p, unsigned v) { p[0] = (unsigned char) v; p[1] = (unsigned char) (v >> 8); }

int fm_wav_header(int mode, unsigned char out[FM_WAV_HEADER_BYTES])
{
    const unsigned ch = mode == 2 ? 2 : 1, rate = 48000, bits = 16;
    const uint32_t byte_rate = rate * ch * bits / 8; /* also the placeholder data size: one second */
    if (!out) return -1;
    memset(out, 0, FM_WAV_HEADER_BYTES);
    memcpy(out, "RIFF", 4);
    le32(out + 4, byte_rate + 36);
    memcpy(out + 8, "WAVEfmt ", 8);
    le32(out + 16, 16);
    le16(out + 20, 1); /* PCM */
    le16(out + 22, ch);
    le32(out + 24, rate);
    le32(out + 28, byte_rate);
    le16(out + 32, ch * bits / 8);
    le16(out + 34, bits);
    memcpy(out + 36, "data", 4);
    le32(out + 40, byte_rate);
    return 0;
}

int fm_wav_open(fm_wav **w, const char *path, int mode)
{
    fm_wav *x;
    unsigned char hdr[FM_WAV_HEADER_BYTES];
    if (!w || !path) return -1;
    *w = NULL;
    x = calloc(1, sizeof *x);
    if (!x) return -1;
    if (strcmp(path, "-") == 0) { x->f = stdout; x->is_stdout = 1; }
    else x->f = fopen(path, "wb");
    if (!x->f) { free(x); return -1; }
    fm_wav_header(mode, hdr);
    if (fwrite(hdr, 1, sizeof hdr, x->f) != sizeof hdr) {
        if (!x->is_stdout) fclose(x->f);
        free(x);
        return -1;
    }
    *w = x;
    return 0;
}

int fm_wav_keep_tail(fm_wav *w, int on)
{
    if (!w) return -1;
    w->keep_tail = on;
    return 0;
}

int fm_wav_write(fm_wav *w, const void *pcm, size_t bytes)
{
    const unsigned char *p = pcm;
    if (!w || (!pcm && bytes)) return -1;
    while (bytes) {
        size_t n = FM_WAV_CLUSTER_BYTES - w->fill;
        if (n > bytes) n = bytes;
        memcpy(w->cluster + w->fill, p, n);
        w->fill += n; p += n; bytes -= n;
        if (w->fill == FM_WAV_CLUSTER_BYTES) {
            if (fwrite(w->cluster, 1, FM_WAV_CLUSTER_BYTES, w->f) != FM_WAV_CLUSTER_BYTES) return -1;
            w->fill = 0;
        }
    }
    return 0;
}

int fm_wav_close(fm_wav *w)
{
    int rc = 0;
    if (!w) return -1;
    if (w->keep_tail && w->fill && fwrite(w->cluster, 1, w->fill, w->f) != w->fill) rc = -1;
    if (!w->is_stdout) {
        unsigned char b[4];
        long size = ftell(w->f);
        if (size < 0) rc = -1;
        else {
            le32(b, (uint32_t) (size - 8));
            if (fseek(w->f, 4, SEEK_SET) != 0 || fwrite(b, 1, 4, w->f) != 4) rc = -1;
            le32(b, (uint32_t) (size - 44));
            if (fseek(w->f, 40, SEEK_SET) != 0 || fwrite(b, 1, 4, w->f) != 4) rc = -1;
        }
        if (fclose(w->f) != 0) rc = -1;
    } else {
        fflush(w->f);
    }
    free(w);
    return rc;
}
