/* oracle/gdstub/gd.h -- TEST INFRASTRUCTURE ONLY.
 * libgd is not in this image.  The reference's bam_sliding_count.c draws a PNG of the hit distribution with libgd
 * (draw_hits, bam_sliding_count.c:276-331) next to the text table that is the parity target (output_count_GC,
 * :140-165).  These no-op declarations let the UNMODIFIED reference source compile and run here; the PNG it
 * "writes" is an empty file and is ignored. */
#ifndef NGSCOV_GDSTUB_H
#define NGSCOV_GDSTUB_H
#include <stdio.h>
typedef struct { int unused; } gdImage;
typedef gdImage *gdImagePtr;
typedef struct { int unused; } gdFont;
typedef gdFont *gdFontPtr;
static gdImage ngscov_gd_image;
static gdFont ngscov_gd_font;
static inline gdImagePtr gdImageCreate(int w, int h) { (void)w; (void)h; return &ngscov_gd_image; }
static inline int gdImageColorExact(gdImagePtr im, int r, int g, int b) { (void)im; (void)r; (void)g; (void)b; return -1; }
static inline void gdImageColorDeallocate(gdImagePtr im, int c) { (void)im; (void)c; }
static inline int gdImageColorAllocate(gdImagePtr im, int r, int g, int b) { (void)im; (void)r; (void)g; (void)b; return 0; }
static inline void gdImageString(gdImagePtr im, gdFontPtr f, int x, int y, unsigned char *s, int c) { (void)im; (void)f; (void)x; (void)y; (void)s; (void)c; }
static inline void gdImageRectangle(gdImagePtr im, int a, int b, int c, int d, int col) { (void)im; (void)a; (void)b; (void)c; (void)d; (void)col; }
static inline void gdImageFilledRectangle(gdImagePtr im, int a, int b, int c, int d, int col) { (void)im; (void)a; (void)b; (void)c; (void)d; (void)col; }
static inline void gdImageSetPixel(gdImagePtr im, int x, int y, int c) { (void)im; (void)x; (void)y; (void)c; }
static inline void gdImagePng(gdImagePtr im, FILE *f) { (void)im; (void)f; }
static inline void gdImageDestroy(gdImagePtr im) { (void)im; }
static inline gdFontPtr gdFontGetGiant(void) { return &ngscov_gd_font; }
#endif
