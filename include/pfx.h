/*
 * pfx.h - OPTIONAL extension entry points of pixelforge-b200.  Nothing here exists in the
 * reference; programs that only include pixelforge.h never need it.  The reference has no flush
 * call because every pfVertex* rasterises synchronously; a batching GPU back end
 * needs one for throughput-oriented callers.
 */
#ifndef PIXEL_FORGE_X_H
#define PIXEL_FORGE_X_H

#include "pixelforge.h"
#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

typedef struct {
    PFuint64 triangles_submitted;    /* Rasterize_Triangle-equivalent invocations sent to the GPU */
    PFuint64 triangles_rasterised;   /* ... surviving the face / zero-area test                   */
    PFuint64 pixels_shaded;          /* colour+depth writes                                       */
    PFuint64 pixels_depth_failed;
    PFuint64 kernel_launches;
    PFuint64 bytes_h2d, bytes_d2h;   /* PCIe traffic caused by this library                          */
} PFXcounters;

/* PF_CUDA_SYNC=end (default): the caller's buffers are refreshed at every pfEnd / pfCallList /
 * pfDraw* / pfClear, as with the reference.  explicit: only at the API's read-back points and at
 * pfxFinish().  pfxSetSyncMode overrides the environment variable. */
PF_API void pfxSetSyncMode(PFboolean explicitSync);
/* Submit everything batched so far on the current context (asynchronous). */
PF_API void pfxFlush(void);
/* pfxFlush + wait + bring the current target's host buffer (and visible z-buffer) up to date. */
PF_API void pfxFinish(void);
PF_API void pfxGetCounters(PFXcounters *out);
PF_API void pfxResetCounters(void);
/* Multi-GPU screen-tile split of the current target: this process renders tiles t with
 * t % world == rank (tile = 64x64 pixels, t = tx + ty*tilesX). */
PF_API void pfxSetTileOwner(PFuint rank, PFuint world);
/* Device pointers of the current target (colour RGBA8 [y*W+x], depth f32) for zero-copy interop. */
PF_API void *pfxGetDeviceColor(void);
PF_API void *pfxGetDeviceDepth(void);
/* Copy of the current target's depth buffer into `out` (width*height floats). */
PF_API void pfxReadDepth(PFfloat *out);
/* Record the triangle/state stream the current context hands to the pfcu C-ABI (it is still
 * rendered).  pfxCaptureEnd flushes and returns pointers to arrays of pfcu_state / pfcu_triangle
 * (include/pfcu.h) that stay valid until the next pfxCaptureBegin or pfDeleteContext. */
PF_API void pfxCaptureBegin(void);
PF_API void pfxCaptureEnd(const void **states, PFuint *nStates, const void **triangles, PFuint *nTriangles);
/* Large PF_TRIANGLES vertex-array draws run their vertex stage (transform, clip, project) on the GPU
 * (default on; PF_CUDA_DEVICE_VERTEX=0 or this call turn it off for the current context). */
PF_API void pfxEnableDeviceVertexStage(PFboolean on);
/* Diagnostics for the Gouraud specular tables the device vertex stage uses instead of powf (pfcu.h,
 * PFCU_POW_TABLE_SIZE): builds (or finds) the table of `shininess` for the current context and compares it with
 * (PFubyte)(255 * powf(x, shininess)) of this host's libm on `samples` pseudo-random x in [0, 1.000001] (the dot product of two normalised vectors) plus both
 * neighbours of every threshold.  Returns the number of mismatches, -1 when the shininess is not tabulated (the
 * host then lights those vertices itself), -2 without a current context. */
PF_API int pfxSpecularTableCheck(PFfloat shininess, PFuint samples);
/* The same kind of check for the fog alpha steps that stand in for the host's expf / exp2f on the device (pfFogProcess in
 * the PF_EXP / PF_EXP2 modes): mismatches on `samples` pseudo-random depths plus the neighbours of every step; -1 when the
 * host function is not monotonic over the fog range, -2 without a context or in PF_LINEAR mode. */
PF_API int pfxFogTableCheck(PFuint samples);
/* Page-locked host memory for buffers the application hands to the library every frame (vertex / index arrays, pixel
 * buffers): copies from it are plain DMA at full PCIe speed instead of being staged by the driver.  Optional: ordinary
 * malloc'ed memory works everywhere, as with the reference.  Free with pfxHostFree (after the last call that used it). */
PF_API void *pfxHostAlloc(size_t bytes);
PF_API void  pfxHostFree(void *p);
/* Static geometry.  The reference reads the vertex and index arrays of pfDrawElements / pfDrawArrays from the caller's
 * memory at every draw; this library copies them to the GPU at every draw, because it cannot know whether the
 * application changed them.  pfxHostStatic(p, PF_TRUE) on a pfxHostAlloc block is the application's promise that it
 * announces changes: the block is copied to the device once and draws whose arrays lie inside it move no vertex data
 * over PCIe; after rewriting (any part of) the block call pfxHostModified(p) and the next draw uploads it again. */
PF_API void pfxHostStatic(void *p, PFboolean isStatic);
PF_API void pfxHostModified(void *p);
/* Texel memory is copied to the device at the first draw that samples the texture; the reference reads the caller's
 * memory at every fragment, so a program that rewrites texels in place (video frames, procedural updates) calls this
 * after each rewrite to have them uploaded again.  Draw calls issued before it keep the old texels. */
PF_API void pfxTextureDirty(PFtexture texture);
/* Explicit sync mode starts the read-back of a context's (page-locked) target buffer as soon as the context is left or
 * its queued work is submitted, so that presenting many contexts does not pay one blocking copy after the other.  A
 * caller that renders contexts it will not read back this frame (device-side consumers, benchmarks of the render rate)
 * turns that off; pfxFinish and the API's read-back points still deliver the pixels. */
PF_API void pfxEnableQueuedReadback(PFboolean on);
/* The pfcu_surface* behind the current target. */
PF_API void *pfxGetSurfaceHandle(void);
PF_API const char *pfxBackendName(void);

#if defined(__cplusplus)
}
#endif
#endif
