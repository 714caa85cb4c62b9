/* fm_timeshift.c -- see include/fm_timeshift.h. */
#include "fm_timeshift.h"

#include <stdlib.h>
#include <string.h>

struct fm_timeshift {
    int slots, n_streams;
    int bottom;   /* circbufferbotton: slot the next cluster is written to */
    int wrapped;  /* circbufferfull */
    int out;      /* circbufferout of the last push */
    char *ring;   /* [n_streams][slots][32768] */
};

int fm_timeshift_slots_for_kbytes(long kbytes)
{
    if (kbytes < 0) return -1;
    return (int) ((kbytes * 1024) / FM_TS_CLUSTER_BYTES);
}

int fm_timeshift_create(fm_timeshift **ts, int slots, int n_streams)
{
    fm_timeshift *t;
    if (!ts || slots < 2 || n_streams < 1) return -1;
    t = (fm_timeshift *) calloc(1, sizeof *t);
    if (!t) return -1;
    t->ring = (char *) malloc((size_t) slots * (size_t) n_streams * FM_TS_CLUSTER_BYTES);
    if (!t->ring) { free(t); return -1; }
    t->slots = slots;
    t->n_streams = n_streams;
    *ts = t;
    return 0;
}

void fm_timeshift_destroy(fm_timeshift *ts)
{
    if (!ts) return;
    free(ts->ring);
    free(ts);
}

static char *slot_of(const fm_timeshift *ts, int stream, int slot)
{
    return ts->ring + ((size_t) stream * (size_t) ts->slots + (size_t) slot) * FM_TS_CLUSTER_BYTES;
}

int fm_timeshift_push(fm_timeshift *ts, const void *clusters, size_t cluster_pitch, int *shift, void *out)
{
    int s, sh;
    if (!ts || !clusters || !shift || cluster_pitch < FM_TS_CLUSTER_BYTES) return -1;
    for (s = 0; s < ts->n_streams; s++)
        memcpy(slot_of(ts, s, ts->bottom), (const char *) clusters + (size_t) s * cluster_pitch, FM_TS_CLUSTER_BYTES);
    sh = *shift;
    if (sh < 0) sh = 0;
    if (!ts->wrapped) {
        if (sh > ts->bottom) sh = ts->bottom;
    } else {
        if (sh > ts->slots - 2) sh = ts->slots - 2;
    }
    ts->out = ts->bottom - sh;
    if (ts->out < 0) ts->out = ts->slots - (sh - ts->bottom);
    *shift = sh;
    if (out)
        for (s = 0; s < ts->n_streams; s++)
            memcpy((char *) out + (size_t) s * FM_TS_CLUSTER_BYTES, slot_of(ts, s, ts->out), FM_TS_CLUSTER_BYTES);
    if (++ts->bottom >= ts->slots) {
        ts->wrapped = 1;
        ts->bottom = 0;
    }
    return ts->out;
}

const void *fm_timeshift_playback(const fm_timeshift *ts, int stream)
{
    if (!ts || stream < 0 || stream >= ts->n_streams) return NULL;
    return slot_of(ts, stream, ts->out);
}

int fm_timeshift_state(const fm_timeshift *ts, int *bottom, int *wrapped, int *slots)
{
    if (!ts) return -1;
    if (bottom) *bottom = ts->bottom;
    if (wrapped) *wrapped = ts->wrapped;
    if (slots) *slots = ts->slots;
    return 0;
}
