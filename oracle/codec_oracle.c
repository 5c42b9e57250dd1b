/*
 * oracle/codec_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Restatement of the decode side of the reference's DDS/PVM codec,
 * /root/reference/src/ddsbase.cpp ("dds:n", third-party V^3 code by S. Roettger):
 *   bit reader dds:132-158, DDS_decode dds:394-452, DDS_interleave dds:167-229,
 *   readDDSfile dds:550-594, readPVMvolume dds:768-858, checksum dds:872-893.
 * Pinned byte-for-byte against the reference's own ddsbase.cpp compiled unmodified into
 * oracle/_ref/ (tests/test_codec.py) and against tests/golden/ fixtures written by it.
 *
 * Deviation: where the reference prints an error and carries on with a malformed file
 * (codebase.h:70-85) this restatement returns NULL.
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DDS_INTERLEAVE_BLOCK (1u << 24)      /* dds:16 */
#define DDS_RL 7                             /* dds:18 */

typedef struct bitreader {
    const uint8_t* cache;
    uint64_t pos, size;      /* size rounded up to a multiple of 4 (dds:128), zero padded */
    uint32_t buffer, bufsize;
} bitreader;

static inline uint32_t shl(uint32_t v, uint32_t b) { return b >= 32 ? 0u : v << b; }
static inline uint32_t shr(uint32_t v, uint32_t b) { return b >= 32 ? 0u : v >> b; }

static uint32_t load_be32(const bitreader* r, uint64_t pos, uint64_t real_size)
{
    uint32_t w = 0;
    for (int i = 0; i < 4; ++i) {
        const uint64_t p = pos + (uint64_t)i;
        const uint32_t byte = p < real_size ? r->cache[p] : 0u;   /* zero padding, dds:125-126 */
        w = (w << 8) | byte;
    }
    return w;
}

static uint32_t readbits(bitreader* r, uint32_t bits, uint64_t real_size)     /* dds:132-158 */
{
    uint32_t value;
    if (bits < r->bufsize) {
        r->bufsize -= bits;
        value = shr(r->buffer, r->bufsize);
    } else {
        value = shl(r->buffer, bits - r->bufsize);
        if (r->pos >= r->size) r->buffer = 0;
        else { r->buffer = load_be32(r, r->pos, real_size); r->pos += 4; }
        r->bufsize += 32 - bits;
        value |= shr(r->buffer, r->bufsize);
    }
    r->buffer &= shl(1u, r->bufsize) - 1u;
    return value;
}

/* inverse of the de-interleave, dds:167-229 with restore=TRUE */
static int interleave_restore(uint8_t* data, uint64_t bytes, uint32_t skip, uint64_t block)
{
    if (skip <= 1) return 0;
    if (block == 0) {
        uint8_t* tmp = (uint8_t*)malloc(bytes ? bytes : 1);
        if (!tmp) return -1;
        const uint8_t* ptr = data;
        for (uint32_t i = 0; i < skip; ++i)
            for (uint64_t j = i; j < bytes; j += skip) tmp[j] = *ptr++;
        memcpy(data, tmp, bytes);
        free(tmp);
        return 0;
    }
    const uint64_t chunk = (uint64_t)skip * block;
    uint8_t* tmp = (uint8_t*)malloc(bytes < chunk ? (bytes ? bytes : 1) : chunk);
    if (!tmp) return -1;
    uint64_t k;
    for (k = 0; k < bytes / skip / block; ++k) {
        const uint8_t* ptr = data + k * chunk;
        for (uint32_t i = 0; i < skip; ++i)
            for (uint64_t j = i; j < chunk; j += skip) tmp[j] = *ptr++;
        memcpy(data + k * chunk, tmp, chunk);
    }
    const uint64_t rest = bytes - k * chunk;
    const uint8_t* ptr = data + k * chunk;
    for (uint32_t i = 0; i < skip; ++i)
        for (uint64_t j = i; j < rest; j += skip) tmp[j] = *ptr++;
    memcpy(data + k * chunk, tmp, rest);
    free(tmp);
    return 0;
}

/* dds:394-452 */
static uint8_t* dds_decode(const uint8_t* chunk, uint64_t size, uint64_t block, uint64_t* out_bytes)
{
    bitreader r;
    r.cache = chunk; r.pos = 0; r.size = 4 * ((size + 3) / 4); r.buffer = 0; r.bufsize = 0;

    const uint32_t skip = readbits(&r, 2, size) + 1;
    const uint32_t strip = readbits(&r, 16, size) + 1;

    uint64_t cap = 1u << 20, cnt = 0;
    uint8_t* out = (uint8_t*)malloc(cap);
    if (!out) return NULL;
    int act = 0;
    uint32_t cnt1;
    while ((cnt1 = readbits(&r, DDS_RL, size)) != 0) {
        const uint32_t code = readbits(&r, 3, size);
        const uint32_t bits = code >= 1 ? code + 1 : code;            /* dds:163-164 */
        for (uint32_t c2 = 0; c2 < cnt1; ++c2) {
            const int delta = (int)readbits(&r, bits, size) - (int)((1u << bits) / 2);
            if (strip == 1 || cnt <= strip) act += delta;
            else act += (int)out[cnt - strip] - (int)out[cnt - strip - 1] + delta;   /* dds:423 */
            while (act < 0) act += 256;
            while (act > 255) act -= 256;
            if (cnt == cap) {
                cap *= 2;
                uint8_t* n = (uint8_t*)realloc(out, cap);
                if (!n) { free(out); return NULL; }
                out = n;
            }
            out[cnt++] = (uint8_t)act;
        }
    }
    if (interleave_restore(out, cnt, skip, block) != 0) { free(out); return NULL; }
    *out_bytes = cnt;
    return out;
}

uint8_t* orc_pvm_decode(const uint8_t* file, uint64_t file_bytes,
                        uint32_t* w, uint32_t* h, uint32_t* d, uint32_t* components,
                        float scale[3], uint64_t* payload_bytes)
{
    static const char ID1[] = "DDS v3d\n", ID2[] = "DDS v3e\n";       /* dds:22-23 */
    uint8_t* data = NULL;
    uint64_t bytes = 0;
    /* readDDSfile dds:550-594, else plain file dds:783-784 */
    if (file_bytes >= 8 && memcmp(file, ID1, 8) == 0)
        data = dds_decode(file + 8, file_bytes - 8, 0, &bytes);
    else if (file_bytes >= 8 && memcmp(file, ID2, 8) == 0)
        data = dds_decode(file + 8, file_bytes - 8, DDS_INTERLEAVE_BLOCK, &bytes);
    else {
        data = (uint8_t*)malloc(file_bytes + 1);
        if (data) { memcpy(data, file, file_bytes); bytes = file_bytes; }
    }
    if (!data) return NULL;
    if (bytes < 5) { free(data); return NULL; }
    {
        uint8_t* n = (uint8_t*)realloc(data, bytes + 1);
        if (!n) { free(data); return NULL; }
        data = n;
        data[bytes] = '\0';
    }

    /* readPVMvolume dds:788-858 */
    int version = 1;
    float sx = 1.0f, sy = 1.0f, sz = 1.0f;
    unsigned int W = 0, H = 0, D = 0, numc = 0;
    char* ptr;
    if (strncmp((char*)data, "PVM\n", 4) != 0) {
        if (strncmp((char*)data, "PVM2\n", 5) == 0) version = 2;
        else if (strncmp((char*)data, "PVM3\n", 5) == 0) version = 3;
        else { free(data); return NULL; }
        ptr = (char*)&data[5];
        if (sscanf(ptr, "%u %u %u\n%g %g %g\n", &W, &H, &D, &sx, &sy, &sz) != 6) { free(data); return NULL; }
        if (W < 1 || H < 1 || D < 1 || sx <= 0.0f || sy <= 0.0f || sz <= 0.0f) { free(data); return NULL; }
        ptr = strchr(ptr, '\n');
        if (!ptr) { free(data); return NULL; }
        ptr += 1;
    } else {
        ptr = (char*)&data[4];
        while (*ptr == '#')
            while (*ptr++ != '\n') { if (ptr >= (char*)data + bytes) { free(data); return NULL; } }
        if (sscanf(ptr, "%u %u %u\n", &W, &H, &D) != 3) { free(data); return NULL; }
        if (W < 1 || H < 1 || D < 1) { free(data); return NULL; }
    }
    ptr = strchr(ptr, '\n');
    if (!ptr) { free(data); return NULL; }
    ptr += 1;
    if (sscanf(ptr, "%u\n", &numc) != 1 || numc < 1) { free(data); return NULL; }
    ptr = strchr(ptr, '\n');
    if (!ptr) { free(data); return NULL; }
    ptr += 1;

    const uint64_t vol = (uint64_t)W * H * D * numc;
    const uint8_t* end = data + bytes;
    if ((const uint8_t*)ptr + vol > end) { free(data); return NULL; }
    uint64_t lens = 0;
    if (version == 3) {
        const uint8_t* q = (const uint8_t*)ptr + vol;
        for (int i = 0; i < 4; ++i) {
            const uint64_t l = strlen((const char*)q) + 1;
            lens += l; q += l;
            if (q > end) { free(data); return NULL; }
        }
    }
    if (end != (const uint8_t*)ptr + vol + lens) { free(data); return NULL; }   /* dds:833 */

    uint8_t* volume = (uint8_t*)malloc(vol + lens ? vol + lens : 1);
    if (!volume) { free(data); return NULL; }
    memcpy(volume, ptr, vol + lens);
    free(data);
    *w = W; *h = H; *d = D; *components = numc;
    scale[0] = sx; scale[1] = sy; scale[2] = sz;
    *payload_bytes = vol + lens;
    return volume;
}

uint32_t orc_checksum(const uint8_t* data, uint64_t bytes)          /* dds:872-893 */
{
    const uint32_t prime = 271;
    uint32_t sum = 0, cipher = 1;
    for (uint64_t i = 0; i < bytes; ++i) {
        const uint8_t value = data[i];
        cipher = prime * cipher + value;
        sum += cipher * value;
    }
    return sum;
}

void orc_free(void* p) { free(p); }
