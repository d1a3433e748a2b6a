/*
 * pixelforge.h - public C API of pixelforge-b200.
 *
 * Drop-in for the header of Bigfoot71/PixelForge (reference: src/pixelforge.h): same 128 entry
 * points, same enum values, same struct layouts, same opaque handles, so a program written against
 * the reference links against libpixelforge.so unchanged.  Everything behind it is new: a C99
 * batching state machine feeding hand-written sm_100a kernels (see include/pfcu.h).
 *
 * Section markers give the reference line ranges each block has to stay ABI-compatible with.
 * No C++ appears in this header.
 */
#ifndef PIXEL_FORGE_H
#define PIXEL_FORGE_H

#include <stdint.h>

/* ---- linkage / storage-class knobs (reference: src/pixelforge.h:27-74) ---------------------- */

#ifndef PF_API
#  if defined(_WIN32) && defined(PF_BUILD_SHARED)
#    define PF_API __declspec(dllexport)
#  elif defined(_WIN32) && defined(USE_LIBTYPE_SHARED)
#    define PF_API __declspec(dllimport)
#  elif defined(PF_BUILD_SHARED)
#    define PF_API __attribute__((visibility("default")))
#  else
#    define PF_API
#  endif
#endif

/* One current context per thread (the reference degrades this to a process global under
 * GCC+OpenMP, src/pixelforge.h:48-64; we always keep it thread-local). */
#ifndef PF_CTX_DECL
#  if defined(_MSC_VER)
#    define PF_CTX_DECL __declspec(thread)
#  else
#    define PF_CTX_DECL __thread
#  endif
#endif

#ifndef PF_RESTRICT
#  ifdef _MSC_VER
#    define PF_RESTRICT __restrict
#  else
#    define PF_RESTRICT restrict
#  endif
#endif

/* ---- overridable allocator + small helpers (reference: src/pixelforge.h:76-128) ------------- */

#ifndef PF_MALLOC
#  define PF_MALLOC(size) malloc(size)
#endif
#ifndef PF_CALLOC
#  define PF_CALLOC(count, size) calloc(count, size)
#endif
#ifndef PF_REALLOC
#  define PF_REALLOC(ptr, newSize) realloc(ptr, newSize)
#endif
#ifndef PF_FREE
#  define PF_FREE(ptr) free(ptr)
#endif

#ifndef PF_INV_255
#  define PF_INV_255 (1.0 / 255)
#endif
#ifndef PF_MIN_255   /* saturate an int to <= 255, branch-free */
#  define PF_MIN_255(n) ((PFubyte)((PFint)(n) | ((255 - (PFint)(n)) >> 31)))
#endif
#ifndef PF_MAX_0     /* saturate an int to >= 0, branch-free */
#  define PF_MAX_0(n) ((PFubyte)((PFint)(n) & -((PFint)(n) >= 0)))
#endif
#ifndef PF_MIN
#  define PF_MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef PF_MAX
#  define PF_MAX(a, b) ((a) > (b) ? (a) : (b))
#endif
#ifndef PF_CLAMP
#  define PF_CLAMP(x, lo, hi) ((x) < (lo) ? (lo) : ((x) > (hi) ? (hi) : (x)))
#endif

/* ---- scalar types (reference: src/pixelforge.h:130-166) ------------------------------------- */

#if (defined(__STDC_VERSION__) && __STDC_VERSION__ >= 199901L) || defined(__cplusplus) || \
    (defined(_MSC_VER) && _MSC_VER >= 1800)
#  ifndef __cplusplus
#    include <stdbool.h>
#  endif
   typedef bool PFboolean;
#  define PF_FALSE 0
#  define PF_TRUE  1
#else
   typedef enum { PF_FALSE = 0, PF_TRUE = 1 } PFboolean;
#endif

typedef int8_t    PFbyte;
typedef uint8_t   PFubyte;
typedef int16_t   PFshort;
typedef uint16_t  PFushort;
typedef int32_t   PFint;
typedef uint32_t  PFuint;
typedef int64_t   PFint64;
typedef uint64_t  PFuint64;
typedef uint32_t  PFsizei;
typedef uint32_t  PFenum;
typedef intptr_t  PFintptr;
typedef uintptr_t PFsizeiptr;
typedef float     PFfloat;
typedef double    PFdouble;

typedef enum {
    PF_UNSIGNED_BYTE = 0, PF_UNSIGNED_SHORT, PF_UNSIGNED_SHORT_5_6_5, PF_UNSIGNED_SHORT_5_5_5_1,
    PF_UNSIGNED_SHORT_4_4_4_4, PF_UNSIGNED_INT, PF_BYTE, PF_SHORT, PF_INT, PF_HALF_FLOAT,
    PF_FLOAT, PF_DOUBLE
} PFdatatype;

/* ---- context enums (reference: src/pixelforge.h:168-271) ------------------------------------ */

typedef void* PFcontext;   /* opaque */

typedef enum {             /* bit flags for pfEnable / pfDisable / pfIsEnabled */
    PF_TEXTURE_2D          = 0x0001,
    PF_FRAMEBUFFER         = 0x0002,
    PF_BLEND               = 0x0004,
    PF_DEPTH_TEST          = 0x0008,
    PF_CULL_FACE           = 0x0010,
    PF_NORMALIZE           = 0x0020,
    PF_LIGHTING            = 0x0040,
    PF_COLOR_MATERIAL      = 0x0080,
    PF_VERTEX_ARRAY        = 0x0100,
    PF_NORMAL_ARRAY        = 0x0200,
    PF_COLOR_ARRAY         = 0x0400,
    PF_TEXTURE_COORD_ARRAY = 0x0800
} PFstate;

typedef enum {             /* pfGet*v names; numbering starts at 10000 and is dense */
    PF_VIEWPORT = 10000,
    PF_COLOR_CLEAR_VALUE,
    PF_DEPTH_CLEAR_VALUE,
    PF_CULL_FACE_MODE,
    PF_CURRENT_COLOR,
    PF_CURRENT_NORMAL,
    PF_CURRENT_TEXTURE_COORDS,
    PF_CURRENT_RASTER_POSITION,
    PF_BLEND_FUNC,
    PF_DEPTH_FUNC,
    PF_POLYGON_MODE,
    PF_POINT_SIZE,
    PF_LINE_WIDTH,
    PF_MATRIX_MODE,
    PF_PROJECTION_MATRIX,
    PF_MODELVIEW_MATRIX,
    PF_TEXTURE_MATRIX,
    PF_MAX_PROJECTION_STACK_DEPTH,
    PF_MAX_MODELVIEW_STACK_DEPTH,
    PF_MAX_TEXTURE_STACK_DEPTH,
    PF_SHADE_MODEL,
    PF_MAX_LIGHTS,
    PF_VERTEX_ARRAY_SIZE,
    PF_VERTEX_ARRAY_STRIDE,
    PF_VERTEX_ARRAY_TYPE,
    PF_NORMAL_ARRAY_STRIDE,
    PF_NORMAL_ARRAY_TYPE,
    PF_TEXTURE_COORD_ARRAY_STRIDE,
    PF_TEXTURE_COORD_ARRAY_TYPE,
    PF_COLOR_ARRAY_SIZE,
    PF_COLOR_ARRAY_STRIDE,
    PF_COLOR_ARRAY_TYPE,
    PF_ZOOM_X,
    PF_ZOOM_Y
} PFgettable;

typedef enum {
    PF_NO_ERROR = 0,
    PF_INVALID_ENUM,
    PF_INVALID_VALUE,
    PF_STACK_OVERFLOW,
    PF_STACK_UNDERFLOW,
    PF_INVALID_OPERATION,
    PF_ERROR_OUT_OF_MEMORY
#ifndef NDEBUG
    , PF_DEBUG_NO_ERROR, PF_DEBUG_INVALID_ENUM, PF_DEBUG_INVALID_VALUE, PF_DEBUG_STACK_OVERFLOW,
    PF_DEBUG_STACK_UNDERFLOW, PF_DEBUG_INVALID_OPERATION, PF_DEBUG_ERROR_OUT_OF_MEMORY
#endif
} PFerrcode;

/* ---- render enums (reference: src/pixelforge.h:273-397) ------------------------------------- */

typedef enum { PF_COLOR_BUFFER_BIT = 0x01, PF_DEPTH_BUFFER_BIT = 0x02 } PFclearflag;
typedef enum { PF_MODELVIEW = 0, PF_PROJECTION, PF_TEXTURE } PFmatrixmode;

typedef enum {
    PF_POINTS = 0, PF_LINES, PF_TRIANGLES, PF_TRIANGLE_FAN, PF_TRIANGLE_STRIP,
    PF_QUADS, PF_QUAD_FAN, PF_QUAD_STRIP
} PFdrawmode;

typedef enum {
    PF_BLEND_AVERAGE = 0, PF_BLEND_ALPHA, PF_BLEND_ADD, PF_BLEND_SUB,
    PF_BLEND_MUL, PF_BLEND_SCREEN, PF_BLEND_LIGHTEN, PF_BLEND_DARKEN
} PFblendmode;

typedef enum { PF_EQUAL = 0, PF_NOTEQUAL, PF_LESS, PF_LEQUAL, PF_GREATER, PF_GEQUAL } PFdepthmode;
typedef enum { PF_POINT = 0, PF_LINE, PF_FILL } PFpolygonmode;
typedef enum { PF_FLAT = 0, PF_SMOOTH } PFshademode;
typedef enum { PF_GOURAUD = 0, PF_PHONG } PFlightmode;
typedef enum { PF_FRONT = 0, PF_BACK = 1, PF_FRONT_AND_BACK } PFface;
typedef enum { PF_REPEAT = 0, PF_MIRRORED_REPEAT, PF_CLAMP_TO_EDGE } PFtexturewrap;
typedef enum { PF_NEAREST = 0, PF_BILINEAR } PFtexturefilter;

typedef enum {
    PF_LIGHT0 = 0, PF_LIGHT1, PF_LIGHT2, PF_LIGHT3, PF_LIGHT4, PF_LIGHT5, PF_LIGHT6, PF_LIGHT7, PF_LIGHT8
} PFlights;

/* colour selectors shared by lights and materials, then the per-kind parameters */
typedef enum { PF_AMBIENT_AND_DIFFUSE = 1, PF_AMBIENT = 2, PF_DIFFUSE = 3, PF_SPECULAR = 4 } PFrendercolor;
typedef enum { PF_EMISSION = 5, PF_SHININESS = 6 } PFmaterialparam;
typedef enum {
    PF_POSITION = 7, PF_SPOT_DIRECTION = 8,
    PF_SPOT_INNER_CUTOFF = 10, PF_SPOT_OUTER_CUTOFF = 11,
    PF_CONSTANT_ATTENUATION = 12, PF_LINEAR_ATTENUATION = 13, PF_QUADRATIC_ATTENUATION = 14
} PFlightparam;

typedef enum { PF_FOG_MODE = 0, PF_FOG_DENSITY, PF_FOG_START, PF_FOG_END, PF_FOG_COLOR } PFfogparam;
typedef enum { PF_LINEAR = 0, PF_EXP, PF_EXP2 } PFfogmode;

/* ---- value types (reference: src/pixelforge.h:399-431) -------------------------------------- */

typedef struct { PFubyte r, g, b, a; } PFcolor;

/* called once per pixel by pfPostProcess: (x, y, depth, colour) -> new colour */
typedef PFcolor (*PFpostprocessfunc)(PFint, PFint, PFfloat, PFcolor);

typedef enum {
    PF_RED = 0, PF_GREEN, PF_BLUE, PF_ALPHA, PF_LUMINANCE, PF_LUMINANCE_ALPHA,
    PF_RGB, PF_RGBA, PF_BGR, PF_BGRA
} PFpixelformat;

typedef void* PFtexture;      /* opaque */
typedef void* PFrenderlist;   /* opaque */

typedef struct {              /* public, returned by value from pfGenFramebuffer */
    PFtexture texture;
    PFfloat  *zbuffer;
} PFframebuffer;

#if defined(__cplusplus)
extern "C" {
#endif

/* ---- context (reference: src/pixelforge.h:452-640, src/context.c:115-375) -------------------- */
PF_API PFcontext pfCreateContext(void* targetBuffer, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type);
PF_API void      pfDeleteContext(PFcontext ctx);
PF_API void      pfSetMainBuffer(void* targetBuffer, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type);
PF_API void      pfSetAuxBuffer(void *auxFramebuffer);
PF_API void      pfSwapBuffers(void);
PF_API PFcontext pfGetCurrentContext(void);
PF_API void      pfMakeCurrent(PFcontext ctx);
PF_API PFboolean pfIsEnabled(PFstate state);
PF_API void      pfEnable(PFstate state);
PF_API void      pfDisable(PFstate state);

/* ---- getters (reference: src/getter.c) ------------------------------------------------------ */
PF_API void      pfGetBooleanv(PFenum pname, PFboolean* params);
PF_API void      pfGetIntegerv(PFenum pname, PFint* params);
PF_API void      pfGetFloatv(PFenum pname, PFfloat* params);
PF_API void      pfGetDoublev(PFenum pname, PFdouble* params);
PF_API void      pfGetPointerv(PFenum pname, const void** params);
PF_API PFerrcode pfGetError(void);

/* ---- matrices (reference: src/context.c:395-548) -------------------------------------------- */
PF_API void pfMatrixMode(PFmatrixmode mode);
PF_API void pfPushMatrix(void);
PF_API void pfPopMatrix(void);
PF_API void pfLoadIdentity(void);
PF_API void pfTranslatef(PFfloat x, PFfloat y, PFfloat z);
PF_API void pfRotatef(PFfloat angle, PFfloat x, PFfloat y, PFfloat z);
PF_API void pfScalef(PFfloat x, PFfloat y, PFfloat z);
PF_API void pfMultMatrixf(const PFfloat* mat);
PF_API void pfFrustum(PFfloat left, PFfloat right, PFfloat bottom, PFfloat top, PFfloat znear, PFfloat zfar);
PF_API void pfOrtho(PFfloat left, PFfloat right, PFfloat bottom, PFfloat top, PFfloat znear, PFfloat zfar);

/* ---- render state (reference: src/context.c:553-797) ---------------------------------------- */
PF_API void pfViewport(PFint x, PFint y, PFsizei width, PFsizei height);
PF_API void pfPolygonMode(PFface face, PFpolygonmode mode);
PF_API void pfShadeModel(PFshademode mode);
PF_API void pfLightModel(PFlightmode mode);
PF_API void pfLineWidth(PFfloat width);
PF_API void pfPointSize(PFfloat size);
PF_API void pfCullFace(PFface face);
PF_API void pfBlendFunc(PFblendmode mode);
PF_API void pfDepthFunc(PFdepthmode mode);
PF_API void pfBindFramebuffer(PFframebuffer* framebuffer);
PF_API void pfBindTexture(PFtexture texture);
PF_API void pfClear(PFclearflag flag);
PF_API void pfClearDepth(PFfloat depth);
PF_API void pfClearColor(PFubyte r, PFubyte g, PFubyte b, PFubyte a);

/* ---- lights and materials (reference: src/context.c:802-1156) ------------------------------- */
PF_API void      pfEnableLight(PFsizei light);
PF_API void      pfDisableLight(PFsizei light);
PF_API PFboolean pfIsEnabledLight(PFsizei light);
PF_API void      pfLightf(PFsizei light, PFenum param, PFfloat value);
PF_API void      pfLightfv(PFsizei light, PFenum param, const void* value);
PF_API void      pfMaterialf(PFface face, PFenum param, PFfloat value);
PF_API void      pfMaterialfv(PFface face, PFenum param, const void* value);
PF_API void      pfColorMaterial(PFface face, PFenum mode);

/* ---- vertex arrays (reference: src/context.c:1161-1575) ------------------------------------- */
PF_API void pfVertexPointer(PFint size, PFenum type, PFsizei stride, const void* pointer);
PF_API void pfNormalPointer(PFenum type, PFsizei stride, const void* pointer);
PF_API void pfTexCoordPointer(PFenum type, PFsizei stride, const void* pointer);
PF_API void pfColorPointer(PFint size, PFenum type, PFsizei stride, const void* pointer);
PF_API void pfDrawElements(PFdrawmode mode, PFsizei count, PFdatatype type, const void* indices);
PF_API void pfDrawArrays(PFdrawmode mode, PFint first, PFsizei count);

/* ---- immediate mode (reference: src/context.c:1580-1913) ------------------------------------ */
PF_API void pfBegin(PFdrawmode mode);
PF_API void pfEnd(void);
PF_API void pfVertex2i(PFint x, PFint y);
PF_API void pfVertex2f(PFfloat x, PFfloat y);
PF_API void pfVertex2fv(const PFfloat* v);
PF_API void pfVertex3i(PFint x, PFint y, PFint z);
PF_API void pfVertex3f(PFfloat x, PFfloat y, PFfloat z);
PF_API void pfVertex3fv(const PFfloat* v);
PF_API void pfVertex4i(PFint x, PFint y, PFint z, PFint w);
PF_API void pfVertex4f(PFfloat x, PFfloat y, PFfloat z, PFfloat w);
PF_API void pfVertex4fv(const PFfloat* v);
PF_API void pfColor(PFcolor color);
PF_API void pfColor1ui(PFuint color);
PF_API void pfColor3ub(PFubyte r, PFubyte g, PFubyte b);
PF_API void pfColor3ubv(const PFubyte* v);
PF_API void pfColor3us(PFushort r, PFushort g, PFushort b);
PF_API void pfColor3usv(const PFushort* v);
PF_API void pfColor3ui(PFuint r, PFuint g, PFuint b);
PF_API void pfColor3uiv(const PFuint* v);
PF_API void pfColor3f(PFfloat r, PFfloat g, PFfloat b);
PF_API void pfColor3fv(const PFfloat* v);
PF_API void pfColor4ub(PFubyte r, PFubyte g, PFubyte b, PFubyte a);
PF_API void pfColor4ubv(const PFubyte* v);
PF_API void pfColor4us(PFushort r, PFushort g, PFushort b, PFushort a);
PF_API void pfColor4usv(const PFushort* v);
PF_API void pfColor4ui(PFuint r, PFuint g, PFuint b, PFuint a);
PF_API void pfColor4uiv(const PFuint* v);
PF_API void pfColor4f(PFfloat r, PFfloat g, PFfloat b, PFfloat a);
PF_API void pfColor4fv(const PFfloat* v);
PF_API void pfTexCoord2f(PFfloat u, PFfloat v);
PF_API void pfTexCoordfv(const PFfloat* v);
PF_API void pfNormal3f(PFfloat x, PFfloat y, PFfloat z);
PF_API void pfNormal3fv(const PFfloat* v);

/* ---- rectangles, pixel blits, fog, misc (reference: src/context.c:1918-2441) ---------------- */
PF_API void pfRects(PFshort x1, PFshort y1, PFshort x2, PFshort y2);
PF_API void pfRectsv(const PFshort* v1, const PFshort* v2);
PF_API void pfRectf(PFfloat x1, PFfloat y1, PFfloat x2, PFfloat y2);
PF_API void pfRectfv(const PFfloat* v1, const PFfloat* v2);
PF_API void pfDrawPixels(PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type, const void* pixels);
PF_API void pfPixelZoom(PFfloat xfactor, PFfloat yfactor);
PF_API void pfRasterPos2i(PFint x, PFint y);
PF_API void pfRasterPos2f(PFfloat x, PFfloat y);
PF_API void pfRasterPos2fv(const PFfloat* v);
PF_API void pfRasterPos3i(PFint x, PFint y, PFint z);
PF_API void pfRasterPos3f(PFfloat x, PFfloat y, PFfloat z);
PF_API void pfRasterPos3fv(const PFfloat* v);
PF_API void pfRasterPos4i(PFint x, PFint y, PFint z, PFint w);
PF_API void pfRasterPos4f(PFfloat x, PFfloat y, PFfloat z, PFfloat w);
PF_API void pfRasterPos4fv(const PFfloat* v);
PF_API void pfFogi(PFfogparam pname, PFint param);
PF_API void pfFogf(PFfogparam pname, PFfloat param);
PF_API void pfFogiv(PFfogparam pname, PFint* param);
PF_API void pfFogfv(PFfogparam pname, PFfloat* param);
PF_API void pfFogProcess(void);
PF_API void pfReadPixels(PFint x, PFint y, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type, void* pixels);
PF_API void pfPostProcess(PFpostprocessfunc postProcessFunction);

/* ---- render lists (reference: src/renderlist.c) --------------------------------------------- */
PF_API PFrenderlist pfGenList(void);
PF_API void         pfDeleteList(PFrenderlist* renderList);
PF_API void         pfNewList(PFrenderlist renderList);
PF_API void         pfEndList(void);
PF_API void         pfCallList(const PFrenderlist renderList);

/* ---- framebuffer objects (reference: src/framebuffer.c) ------------------------------------- */
PF_API PFframebuffer pfGenFramebuffer(PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type);
PF_API void          pfDeleteFramebuffer(PFframebuffer* framebuffer);
PF_API PFboolean     pfIsValidFramebuffer(PFframebuffer* framebuffer);
PF_API void          pfClearFramebuffer(PFframebuffer* framebuffer, PFcolor color, PFfloat depth);
PF_API PFcolor       pfGetFramebufferPixel(const PFframebuffer* framebuffer, PFsizei x, PFsizei y);
PF_API PFfloat       pfGetFramebufferDepth(const PFframebuffer* framebuffer, PFsizei x, PFsizei y);
PF_API void          pfSetFramebufferPixelDepthTest(PFframebuffer* framebuffer, PFsizei x, PFsizei y, PFfloat z, PFcolor color, PFdepthmode depthMode);
PF_API void          pfSetFramebufferPixelDepth(PFframebuffer* framebuffer, PFsizei x, PFsizei y, PFfloat z, PFcolor color);
PF_API void          pfSetFramebufferPixel(PFframebuffer* framebuffer, PFsizei x, PFsizei y, PFcolor color);

/* ---- texture objects (reference: src/texture.c) --------------------------------------------- */
PF_API PFtexture pfGenTexture(void* pixels, PFsizei width, PFsizei height, PFpixelformat format, PFdatatype type);
PF_API void      pfDeleteTexture(PFtexture* texture, PFboolean freeBuffer);
PF_API PFboolean pfIsValidTexture(const PFtexture texture);
PF_API void      pfTextureParameter(PFtexture texture, PFtexturewrap wrapMode, PFtexturefilter filterMode);
PF_API void*     pfGetTexturePixels(const PFtexture texture, PFsizei* width, PFsizei* height, PFpixelformat* format, PFdatatype* type);

#if defined(__cplusplus)
}
#endif

#endif /* PIXEL_FORGE_H */
