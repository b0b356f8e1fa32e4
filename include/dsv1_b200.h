/*
 * dsv1_b200.h -- public C ABI of libdsv1_b200.so (B200 / sm_100a implementation of the
 * DSV1 per-frame encode/decode hot path).
 *
 * This one header declares everything a caller of the reference's dsv.h,
 * dsv_encoder.h and dsv_decoder.h can see: same type names, same field names,
 * same field order and types (so struct layouts are byte-identical -- checked by
 * tests/test_abi.py against the reference headers), same function names,
 * argument meaning and error behaviour.  include/compat/{dsv,dsv_encoder,
 * dsv_decoder,util}.h forward to it so that the reference CLI (dsv_main.c)
 * compiles unmodified against this library.
 *
 * Each declaration cites the reference interface it replaces (file:line under
 * the reference tree).  Pixel/coefficient/bit work behind these entry points
 * runs in hand-written CUDA kernels; there is no CPU fallback.
 */
#ifndef DSV1_B200_H
#define DSV1_B200_H

#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* Bitstream constants (dsv.h:26-51)                                          */
/* ------------------------------------------------------------------------- */
#define DSV_FOURCC_0 'D'
#define DSV_FOURCC_1 'S'
#define DSV_FOURCC_2 'V'
#define DSV_FOURCC_3 '1'
#define DSV_VERSION_MINOR 0

#define DSV_PT_META 0x00
#define DSV_PT_PIC 0x04
#define DSV_PT_EOS 0x10
#define DSV_MAKE_PT(is_ref, has_ref) (DSV_PT_PIC | ((is_ref) << 1) | (has_ref))
#define DSV_PT_IS_PIC(x) ((x) & 0x4)
#define DSV_PT_IS_REF(x) (((x) & 0x6) == 0x6)
#define DSV_PT_HAS_REF(x) ((x) & 0x1)

#define DSV_PACKET_HDR_SIZE 14 /* fourcc[4] minor[1] type[1] prev[4] next[4] */
#define DSV_PACKET_TYPE_OFFSET 5
#define DSV_PACKET_PREV_OFFSET 6
#define DSV_PACKET_NEXT_OFFSET 10

#define DSV_MIN_BLOCK_SIZE 16
#define DSV_MAX_BLOCK_SIZE 64

/* ------------------------------------------------------------------------- */
/* Small helpers callers rely on (dsv.h:53-64)                                */
/* ------------------------------------------------------------------------- */
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif
#ifndef CLAMP
#define CLAMP(x, a, b) ((x) < (a) ? (a) : ((x) > (b) ? (b) : (x)))
#endif
#define DSV_ROUND_SHIFT(x, shift) (((x) + (1 << (shift)) - 1) >> (shift))
#define DSV_ROUND_POW2(x, pwr) (((x) + (1 << (pwr)) - 1) & ((unsigned) (~0) << (pwr)))
#define DSV_DIV_ROUND(a, b) (((a) + (b) - 1) / (b))

/* chroma subsampling codes (dsv.h:66-82): low 2 bits = vertical shift, next 2 = horizontal */
#define DSV_FMT_FULL_V 0x0
#define DSV_FMT_DIV2_V 0x1
#define DSV_FMT_DIV4_V 0x2
#define DSV_FMT_FULL_H 0x0
#define DSV_FMT_DIV2_H 0x4
#define DSV_FMT_DIV4_H 0x8
#define DSV_SUBSAMP_444 (DSV_FMT_FULL_H | DSV_FMT_FULL_V)
#define DSV_SUBSAMP_422 (DSV_FMT_DIV2_H | DSV_FMT_FULL_V)
#define DSV_SUBSAMP_420 (DSV_FMT_DIV2_H | DSV_FMT_DIV2_V)
#define DSV_SUBSAMP_411 (DSV_FMT_DIV4_H | DSV_FMT_FULL_V)
#define DSV_FORMAT_H_SHIFT(format) (((format) >> 2) & 0x3)
#define DSV_FORMAT_V_SHIFT(format) ((format) & 0x3)

#define DSV_MAX_QP_BITS 11
#define DSV_MAX_QUALITY ((1 << DSV_MAX_QP_BITS) - 1)
#define DSV_QUALITY_PERCENT(pct) (DSV_MAX_QUALITY * (pct) / 100)

/* ------------------------------------------------------------------------- */
/* Data types (dsv.h:84-150, 179-196)                                         */
/* ------------------------------------------------------------------------- */
typedef uint32_t DSV_FNUM;

typedef struct {
    int width, height;
    int subsamp;
    int fps_num, fps_den;
    int aspect_num, aspect_den;
} DSV_META;

/* host-visible 8-bit plane; data points at sample (0,0) (inside the border, if any) */
typedef struct {
    uint8_t *data;
    int len;
    int format;
    int stride;
    int w, h;
    int hs, vs;
} DSV_PLANE;

typedef int32_t DSV_SBC;
typedef struct {
    DSV_SBC *data;
    int width, height;
} DSV_COEFS;

typedef struct {
    uint8_t *alloc; /* NULL for frames wrapping caller memory */
    DSV_PLANE planes[3];
    int refcount;
    int format;
    int width, height;
    int border; /* 0 or 1 (=> DSV_MAX_BLOCK_SIZE samples on every side) */
} DSV_FRAME;

#define DSV_MODE_INTER 0
#define DSV_MODE_INTRA 1
#define DSV_MASK_INTRA00 1
#define DSV_MASK_INTRA01 2
#define DSV_MASK_INTRA10 4
#define DSV_MASK_INTRA11 8
#define DSV_MASK_ALL_INTRA 15

/* one motion block; vector in half-pel units (dsv.h:137-150) */
typedef struct {
    union {
        struct {
            int16_t x, y;
        } mv;
        int32_t all;
    } u;
    uint8_t mode;
    uint8_t submask;
    uint8_t lo_var;
    uint8_t lo_tex;
    uint8_t high_detail;
} DSV_MV;

#define DSV_GET_LINE(p, y) ((p)->data + (y) * (p)->stride)
#define DSV_GET_XY(p, x, y) ((p)->data + (x) + (y) * (p)->stride)

typedef struct {
    DSV_META *vidmeta;
    int is_ref, has_ref;
    int blk_w, blk_h;
    int nblocks_h, nblocks_v;
} DSV_PARAMS;

typedef struct {
    unsigned char *data;
    unsigned len;
} DSV_BUF;

/* ------------------------------------------------------------------------- */
/* Support functions the CLI links (dsv.h:160-213)                            */
/* ------------------------------------------------------------------------- */
void *dsv_alloc(int size);                      /* zero-filled (dsv.c:47-57) */
void dsv_free(void *ptr);                       /* dsv.c:59-67 */
void dsv_memory_report(void);                   /* dsv.c:69-77 */
void dsv_mk_buf(DSV_BUF *buf, int size);        /* dsv.c:181-187 */
void dsv_buf_free(DSV_BUF *buffer);             /* dsv.c:172-179 */
int dsv_yuv_write(FILE *out, int fno, DSV_PLANE *fd);                              /* dsv.c:98-129 */
int dsv_yuv_read(FILE *in, int fno, uint8_t *o, int w, int h, int subsamp);        /* dsv.c:131-170 */

DSV_FRAME *dsv_mk_frame(int format, int width, int height, int border);           /* frame.c:63-120 */
DSV_FRAME *dsv_load_planar_frame(int format, void *data, int width, int height);  /* frame.c:122-164 */
DSV_FRAME *dsv_frame_ref_inc(DSV_FRAME *frame);                                   /* frame.c:177-183 */
void dsv_frame_ref_dec(DSV_FRAME *frame);                                         /* frame.c:185-197 */
void dsv_frame_copy(DSV_FRAME *dst, DSV_FRAME *src);                              /* frame.c:199-221 */
DSV_FRAME *dsv_clone_frame(DSV_FRAME *f, int border);                             /* frame.c:166-175 */
DSV_FRAME *dsv_extend_frame(DSV_FRAME *frame);                                    /* frame.c:263-295 */
void dsv_frame_add(DSV_FRAME *dst, DSV_FRAME *src);                               /* bmc.c:304-316 */
int dsv_frame_avg_luma(DSV_FRAME *frame);                                         /* frame.c:223-238 */
void dsv_ds2x_frame_luma(DSV_FRAME *dest, DSV_FRAME *src);                        /* frame.c:240-261 */
DSV_FRAME *dsv_extend_frame_luma(DSV_FRAME *frame);                               /* frame.c:297-327 */
void dsv_plane_xy(DSV_FRAME *f, DSV_PLANE *out, int c, int x, int y);             /* frame.c:329-342 */
void dsv_mk_coefs(DSV_COEFS *c, int format, int width, int height);               /* frame.c:29-61 */

/* logging (dsv.h:215-247, dsv.c:19-39) */
#define DSV_LEVEL_NONE 0
#define DSV_LEVEL_ERROR 1
#define DSV_LEVEL_WARNING 2
#define DSV_LEVEL_INFO 3
#define DSV_LEVEL_DEBUG 4
extern char *dsv_lvlname[DSV_LEVEL_DEBUG + 1];
void dsv_set_log_level(int level);
int dsv_get_log_level(void);

#define DSV_LOG_LVL(level, x)                                        \
    do {                                                             \
        if ((level) <= dsv_get_log_level()) {                        \
            printf("[DSV][%s] ", dsv_lvlname[level]);                \
            printf("%s: %s(%d): ", __FILE__, __FUNCTION__, __LINE__); \
            printf x;                                                \
            printf("\n");                                            \
        }                                                            \
    } while (0)
#define DSV_ERROR(x) DSV_LOG_LVL(DSV_LEVEL_ERROR, x)
#define DSV_WARNING(x) DSV_LOG_LVL(DSV_LEVEL_WARNING, x)
#define DSV_INFO(x) DSV_LOG_LVL(DSV_LEVEL_INFO, x)
#define DSV_DEBUG(x) DSV_LOG_LVL(DSV_LEVEL_DEBUG, x)
#define DSV_ASSERT(x)                      \
    do {                                   \
        if (!(x)) {                        \
            DSV_ERROR(("assert: " #x));    \
            exit(-1);                      \
        }                                  \
    } while (0)

/* ------------------------------------------------------------------------- */
/* Encoder (dsv_encoder.h:27-121)                                             */
/* ------------------------------------------------------------------------- */
#define DSV_GOP_INTRA 0
#define DSV_GOP_INF INT_MAX
#define DSV_ENC_NUM_BUFS 0x03
#define DSV_ENC_FINISHED 0x04
#define DSV_RATE_CONTROL_CRF 0
#define DSV_RATE_CONTROL_ABR 1
#define DSV_MAX_PYRAMID_LEVELS 5
#define DSV_BPF_RESET 256

/* In this implementation the per-frame encoder record is device-side state;
 * callers only ever see the pointer. */
typedef struct _DSV_ENCDATA DSV_ENCDATA;

typedef struct {
    /* -- caller-configurable, written directly as the CLI does (dsv_main.c:463-489) -- */
    int quality; /* 0..DSV_MAX_QUALITY */
    int gop;
    int do_scd;
    int rc_mode;
    int rc_high_motion_nudge;
    unsigned bitrate;
    int max_q_step;
    int min_quality;
    int max_quality;
    int min_I_frame_quality;
    int intra_pct_thresh;
    int scene_change_delta;
    unsigned stable_refresh;
    int pyramid_levels;

    /* -- internal; same slots as the reference so sizeof/offsetof agree -- */
    unsigned rc_quant;
    unsigned bpf_total;
    unsigned bpf_reset;
    int bpf_avg;
    int total_P_frame_q;
    int avg_P_frame_q;
    int last_P_frame_over;
    int back_into_range;

    DSV_FNUM next_fnum;
    DSV_ENCDATA *ref; /* here: the encoder's device context (created on first use) */
    DSV_META vidmeta;
    int prev_link;
    int force_metadata;

    struct DSV_STAB_ACC {
        signed x : 16;
        signed y : 16;
    } *stability;
    unsigned refresh_ctr;
    unsigned char *stable_blocks;

    DSV_FNUM prev_gop;
    int prev_avg_luma;
} DSV_ENCODER;

void dsv_enc_init(DSV_ENCODER *enc);                              /* dsv_encoder.c:696-722 */
void dsv_enc_free(DSV_ENCODER *enc);                              /* dsv_encoder.c:736-751 */
void dsv_enc_set_metadata(DSV_ENCODER *enc, DSV_META *md);        /* dsv_encoder.c:753-757 */
void dsv_enc_force_metadata(DSV_ENCODER *enc);                    /* dsv_encoder.c:759-763 */
void dsv_enc_start(DSV_ENCODER *enc);                             /* dsv_encoder.c:724-734 */
/* Takes the frame reference; returns the number of buffers written to bufs
 * (1, or 2 with the metadata packet first).  dsv_encoder.c:780-854 */
int dsv_enc(DSV_ENCODER *enc, DSV_FRAME *frame, DSV_BUF *bufs);
void dsv_enc_end_of_stream(DSV_ENCODER *enc, DSV_BUF *bufs);      /* dsv_encoder.c:765-778 */

/* "used internally" by the reference encoder but exported by it (dsv_encoder.h:122-132, hme.c:730-741):
 * hierarchical motion estimation over caller-built pyramids.  src[i] / ref[i] are level-i frames WITH borders
 * (level 0 = full size; only luma matters above level 0), levels = index of the coarsest level.  On return
 * mvf[i] (i = 0..levels) are dsv_alloc'd arrays of nblocks_h * nblocks_v vectors the caller frees with dsv_free;
 * the value is the percentage of intra blocks at level 0.  Here the frames are copied to the GPU, searched by the
 * same kernels the encoder uses, and the fields copied back. */
typedef struct {
    DSV_PARAMS *params;
    DSV_FRAME *src[DSV_MAX_PYRAMID_LEVELS + 1];
    DSV_FRAME *ref[DSV_MAX_PYRAMID_LEVELS + 1];
    DSV_MV *mvf[DSV_MAX_PYRAMID_LEVELS + 1];
    int levels;
} DSV_HME;
int dsv_hme(DSV_HME *hme);

/* ------------------------------------------------------------------------- */
/* Decoder (dsv_decoder.h:26-59)                                              */
/* ------------------------------------------------------------------------- */
typedef struct _DSV_IMAGE DSV_IMAGE; /* here: the decoder's device context */

#define DSV_DRAW_STABHQ 1
#define DSV_DRAW_MOVECS 2
#define DSV_DRAW_IBLOCK 4

typedef struct {
    DSV_META vidmeta;
    DSV_IMAGE *ref;
    int draw_info; /* DSV_DRAW_* bits: debug overlay painted on the returned P pictures (dsv_decoder.c:441-447) */
    int got_metadata;
} DSV_DECODER;

#define DSV_DEC_OK 0
#define DSV_DEC_ERROR 1
#define DSV_DEC_EOS 2
#define DSV_DEC_GOT_META 3
#define DSV_DEC_NEED_NEXT 4

/* Consumes (frees) buffer; on DSV_DEC_OK with a picture, *out carries one
 * reference the caller drops with dsv_frame_ref_dec.  dsv_decoder.c:286-472 */
int dsv_dec(DSV_DECODER *d, DSV_BUF *buffer, DSV_FRAME **out, DSV_FNUM *fn);
DSV_META *dsv_get_metadata(DSV_DECODER *d);                       /* dsv_decoder.c:275-284 */
void dsv_dec_free(DSV_DECODER *d);                                /* dsv_decoder.c:267-273 */

/* CLI helpers (util.h / util.c:21-93) -- host-only, kept for link compatibility */
unsigned estimate_bitrate(int quality, int gop, DSV_META *md);
void conv444to422(DSV_PLANE *srcf, DSV_PLANE *dstf);
void conv422to420(DSV_PLANE *srcf, DSV_PLANE *dstf);

#ifdef __cplusplus
}
#endif

#endif /* DSV1_B200_H */
