/* params.h — compile-time DEFAULTS of the run-time switches, under the names the reference uses
 * (optixHello/params.h:24-32). In the reference these macros select code at compile time; here they only
 * seed rdc_ingest_options / rdc_frame_params and the OptixHello command line, every one of them can be
 * overridden at run time. */
#ifndef RDC_PARAMS_H
#define RDC_PARAMS_H

/* Are we loading Orzan-2008 diffusion-curve saves? Swaps the x and y axes, mirrors y and swaps the R and B
 * channels (reference params.h:24). */
#ifndef USE_DIFFUSION_CURVE_SAVE
#define USE_DIFFUSION_CURVE_SAVE true
#endif

/* blur pass, per-ray jitter, OptiX denoiser (reference params.h:27-29). The denoiser is a closed neural
 * model of the OptiX SDK; the switch is accepted and ignored. */
#ifndef USE_BLUR
#define USE_BLUR true
#endif
#ifndef USE_AA
#define USE_AA true
#endif
#ifndef USE_DENOISER
#define USE_DENOISER false
#endif

/* how many portals one ray may pass, at most 31 (reference params.h:31-32) */
#ifndef MAX_TRACE_DEPTH
#define MAX_TRACE_DEPTH 2
#endif

/* the knobs of optixHello.cpp:89-98 */
#define RDC_DEFAULT_ZOOM_FACTOR 1.0f
#define RDC_DEFAULT_OFFSET_X 0.0f
#define RDC_DEFAULT_OFFSET_Y 0.0f
#define RDC_DEFAULT_WEIGHT_DEGREE 0.5f
#define RDC_DEFAULT_CURVE_WIDTH 1e-3f
#define RDC_DEFAULT_ENDCAP_SIZE 8.0f

#endif
