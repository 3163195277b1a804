/* psinfer.h -- C ABI of the B200-native pictorial-structures inference library (libpsinfer.so).
 *
 * The reference (leonid-pishchulin/partapp) has no plugin/FFI layer; the seam is cut at the call
 *
 *   object_detect::computeRootPosteriorRot(part_app, log_part_detections, root_part_posterior,
 *                                          rootpart_idx, joints, flip, bIsSparse, imgidx,
 *                                          best_part_hyp, bSaveMarginals)
 *   -- declared src/libs/libPictStruct/objectdetect.h:269-274,
 *      defined  src/libs/libPictStruct/objectdetect_findrot.cpp:470-727,
 *      called   objectdetect_findrot.cpp:992, objectdetect_roi.cpp:262
 *
 * with computeRotJointMarginal (objectdetect.h:277-282, findrot.cpp:292-456) as the finer test seam.
 * Every entry point below names the reference interface it replaces.  Conventions kept from the
 * reference: grids are contiguous C-order fp32 [rotation][y][x]; unaries are log-domain with
 * LOG_ZERO = -1e6 (libBoostMath/boost_math.h:23) marking unevaluated cells; joints arrive 0-based and
 * already flipped; geometry is double precision; ExpParam rotation/scale ranges are *floats*.
 * Conventions that differ: no assert/abort, no stdout -- every call returns an int status and the
 * message is available from ps_last_error(); no exceptions cross the ABI.
 *
 * Threading: one ps_ctx per (host thread, GPU).  A ctx is not thread-safe; distinct ctxs are fully
 * independent (images shard across GPUs/ctxs with no collective).
 *
 * There is no CPU fallback: ps_create fails with PS_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef PSINFER_H_
#define PSINFER_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_MAX_PARTS 64
#define PS_HYP_VEC 7 /* PartHyp::toVect(): scaleidx, scale, rotidx, rot_deg, x, y, score (objectdetect.h:139-160) */

enum ps_status {
  PS_OK = 0,
  PS_ERR_INVALID = 1,     /* bad argument / violated precondition (an assert in the reference) */
  PS_ERR_CUDA = 2,        /* CUDA runtime failure, or no usable device */
  PS_ERR_STATE = 3,       /* call order violated (e.g. ps_infer before ps_set_joints) */
  PS_ERR_UNSUPPORTED = 4  /* valid in the reference but not implemented here; never silently approximated */
};

enum ps_mem_kind { PS_MEM_HOST = 0, PS_MEM_DEVICE = 1 };

/* Joint::POS_GAUSSIAN / ROT_GAUSSIAN, objectdetect.h:55 */
enum ps_joint_type { PS_JOINT_POS_GAUSSIAN = 1, PS_JOINT_ROT_GAUSSIAN = 2 };

/* flags for ps_infer */
enum ps_infer_flags {
  PS_INFER_SPARSE = 1,        /* bIsSparse of computeRootPosteriorRot (findrot.cpp:740 passes true) */
  PS_INFER_LOCAL_MAX = 2,     /* also extract <=K local maxima per part (findLocalMax, aux.cpp:193-261) */
  PS_INFER_ROOT_HYPS = 4,     /* also extract <=1000 local maxima of the root posterior (findrot.cpp:1037-1038) */
  PS_INFER_KEEP_UNARIES = 8,  /* restore the unaries after the call (what findrot.cpp:847,1001 does around it) */
  PS_INFER_NO_BORDER_STRIP = 16 /* skip the root border strip: the message passing of libDiscPS's
                                   partSampleWithPriorHelper (disc_sample_with_prior.cpp:64-330), a copy of
                                   computeRootPosteriorRot + computePartMarginals that keeps the upright masking but has no
                                   strip; its samples are then drawn from ps_get_marginal(part, scaleidx) */
};

/* The ExpParam / PartConfig fields the path reads:
 * ExpParam.proto:189-195 (scale/rotation ranges), :229 strip_border_detections, :282 roi_save_num_samples;
 * PartConfig.part[i].{is_detect,is_upright,is_root}. */
typedef struct ps_config {
  int device;                 /* CUDA device ordinal */
  int num_parts;              /* part_conf.part_size() */
  int num_rotation_steps;     /* ExpParam.num_rotation_steps */
  float min_part_rotation;    /* degrees */
  float max_part_rotation;
  int num_scale_steps;        /* ExpParam.num_scale_steps */
  float min_object_scale;
  float max_object_scale;
  int height, width;          /* image (= state grid) size */
  int root_idx;               /* 0-based root part; -1: the unique part with is_detect && is_root (findrot.cpp:783-788) */
  unsigned char is_detect[PS_MAX_PARTS];
  unsigned char is_upright[PS_MAX_PARTS];
  unsigned char is_root[PS_MAX_PARTS];
  float strip_border_detections;
  int roi_save_num_samples;   /* K of findLocalMax; ExpParam default 1000 */
  int keep_all_scales;        /* 1: keep the marginals of every scale resident (bSaveMarginals use); 0: last scale only */
  int interpolate;            /* ExpParam.interpolate: ps_set_unary_compact resamples with TM_BILINEAR instead of TM_DIRECT
                               * (partapp.cpp:889-894) */
  int fast_math;              /* 0 (default): parity arithmetic -- every filter tap rounds its product and its sum
                               * separately, in the reference's order: results are bit-identical to the CPU path.
                               * 1: the tap is one fused multiply-add (half the fp32-pipe time of the Gaussian and
                               * rotation filters); argmax records stay identical on the test sets, marginals agree
                               * to ~1e-6 relative (north-star bound: 1e-4).  exp/log, resampling, index and order
                               * rules are the same in both modes. */
} ps_config;

/* object_detect::Joint (objectdetect.h:54-86) after loadJoints (aux.cpp:54-141): 0-based ids, flipped. */
typedef struct ps_joint {
  int type;                   /* ps_joint_type; only ROT_GAUSSIAN is on this path (findrot.cpp:766) */
  int child_idx, parent_idx;
  double offset_c[2];         /* parent_pos - child_pos, pixels */
  double offset_p[2];         /* child_pos - parent_pos, pixels */
  double C[4];                /* 2x2 covariance, row-major */
  double rot_mean, rot_sigma; /* radians */
} ps_joint;

typedef struct ps_ctx ps_ctx;

/* ---- lifetime --------------------------------------------------------------------------------- */

/* Allocates every tree level (unaries, beliefs, root messages, scratch) in HBM once. */
int ps_create(const ps_config *cfg, ps_ctx **out);
void ps_destroy(ps_ctx *ctx);
/* Message of the last failing call on this ctx (ctx == NULL: of the last failing ps_create). */
const char *ps_last_error(const ps_ctx *ctx);
/* Run the ctx's work on an existing CUDA stream (a cudaStream_t passed as void*); NULL = the ctx's own stream. */
int ps_set_stream(ps_ctx *ctx, void *cuda_stream);
int ps_synchronize(ps_ctx *ctx);

/* ---- model ------------------------------------------------------------------------------------ */

/* Replaces the `std::vector<Joint> joints` argument of computeRootPosteriorRot (objectdetect.h:269-274).
 * Validates the topology the reference asserts (root with chains: findrot.cpp:207-210) and precomputes, per
 * joint x direction x scale, the shift tables, Gaussian taps, eigen-frames and scatter maps. */
int ps_set_joints(ps_ctx *ctx, const ps_joint *joints, int num_joints);

/* loadJoints' flip branch (aux.cpp:102-119) as a pure function on one joint: C <- T C T, offsets x-negated,
 * rot_mean negated. */
void ps_flip_joint(ps_joint *joint);

/* partapp_aux.hpp:45-58,123-129,86-92,94-100 */
double ps_rot_from_index(const ps_config *cfg, int rotidx);
double ps_scale_from_index(const ps_config *cfg, int scaleidx);
int ps_index_from_rot(const ps_config *cfg, double rot_deg);

/* ---- unaries: `log_part_detections[part][scale]` ------------------------------------------------ */

/* Copies one [R][H][W] grid into the resident unary of (part, scale).  raw_scores != 0: `src` holds
 * classifier scores as loadScoreGrid returns them (partapp.cpp:830-903) and the device applies
 * clip_scores_fill + computeLogGrid (findrot.cpp:834-845; aux.hpp:42-59; op.hpp:154-167). */
int ps_set_unary(ps_ctx *ctx, int part, int scale, const float *src, int mem_kind, int raw_scores);
/* PartApp::loadScoreGrid (libPartApp/partapp.cpp:830-903) + the unary prep, on the device: `cells` is the compact
 * detector grid `cell_scoregrid{scale,rot}` of one (part, scale), [R][grid_h][grid_w] fp32 with 0 = not evaluated
 * (part_detect::NO_CLASS_VALUE); `Tig` is [R][3][3] row-major doubles, Tig = Ti2 * T2g (partapp.cpp:881-887).
 * Every evaluated cell is scattered with TM_DIRECT semantics (transform.hpp:167-192: x outer, y inner, last writer
 * wins) -- or, with ps_config.interpolate, every image cell gathers with TM_BILINEAR through inverse(Tig)
 * (transform.hpp:196-238) -- then clip_scores_fill + computeLogGrid are applied.  Uploading compact grids instead of image-size grids cuts
 * the host-to-device traffic by the detector stride squared (16x for the shipped configuration, README.md:88). */
int ps_set_unary_compact(ps_ctx *ctx, int part, int scale, const float *cells, int grid_h, int grid_w,
                         const double *Tig, int mem_kind);
/* The same mapping stopped after clip_scores_fill: the resident grid holds detector scores (negatives clipped to 1e-4,
 * unevaluated cells 0), which is what findObjectRoiHelper extracts its detection maxima from before it takes the
 * logarithm (objectdetect_roi.cpp:205-236).  Follow with ps_unary_local_max and ps_log_unary; ps_infer on a raw
 * grid is the caller's error. */
int ps_set_unary_compact_raw(ps_ctx *ctx, int part, int scale, const float *cells, int grid_h, int grid_w,
                             const double *Tig, int mem_kind);
/* The same for n (part, scale) grids that share one detector lattice (gh x gw cells, one Tig[R][3][3]) -- every part of
 * an image, typically: one fill and one scatter launch for all of them instead of two per part.  cells[i] is the
 * compact grid [R][gh][gw] of (parts[i], scales[i]).  Falls back to n single-grid calls when the lattice is not
 * collision-free or interpolate is set (same results either way). */
int ps_set_unaries_compact(ps_ctx *ctx, int n, const int *parts, const int *scales, const float *const *cells, int gh,
                           int gw, const double *Tig, int mem_kind);
/* multi_array_op::computeLogGrid (multi_array_op.hpp:154-167) in place on the resident grid of (part, scale)
 * (objectdetect_roi.cpp:240-242). */
int ps_log_unary(ps_ctx *ctx, int part, int scale);
/* findLocalMax (aux.cpp:193-261) on the resident grid of (part, scale): rows of (rotidx, x, y, score), at most max_n
 * (objectdetect_roi.cpp:226-228). */
int ps_unary_local_max(ps_ctx *ctx, int part, int scale, int max_n, float *out, int *count);
/* Reads the resident (possibly masked) unary back. */
int ps_get_unary(ps_ctx *ctx, int part, int scale, float *dst, int mem_kind);

/* addExtraUnary (icps.cpp:526-548) for grids that are a broadcast of a small table:
 *   table_kind 0: table[R]     (getRotScoreGrid, icps.cpp:228-281)   unary += weight * table[r]
 *   table_kind 1: table[H][W]  (getPosScoreGrid, icps.cpp:366-423)   unary += weight * table[y][x]
 *   table_kind 2: table[H][W]  (setTorsoPosPrior, icps.cpp:183-190)  unary += table[y][x]
 * applied to every scale of `part`. */
int ps_add_unary_table(ps_ctx *ctx, int part, const float *table, int table_kind, float weight);
/* Up to 4 such tables applied to `part` in one pass over its grids, in array order -- the reference adds the rotation
 * score, then the position score, then the torso prior (findrot.cpp:913-949), and every add rounds, so the order is
 * part of the result.  mem_kind says where the tables live (PS_MEM_HOST tables are staged in stream order; the call
 * does not synchronise). */
int ps_add_unary_tables(ps_ctx *ctx, int part, int n, const float *const *tables, const int *table_kinds,
                        const float *weights, int mem_kind);

/* DPM score fusion for full grids `grid[num_rot][H][W]` (num_rot = num_rotation_steps, or 1 = the same grid for every
 * rotation), applied to every scale of `part`:
 *   mode 0: addDPMScore     (icps.cpp:488-524)  unary += weight * grid              (grid already log-domain)
 *   mode 1: addLoadDPMScore (icps.cpp:445-486)  unary += grid > 1e-4 ? weight * log(grid) : log(1e-4) */
int ps_add_unary_grid(ps_ctx *ctx, int part, const float *grid, int num_rot, int mode, float weight, int mem_kind);

/* Host builders of those tables (same arithmetic as the reference, which builds them on the CPU). */
void ps_rot_score_table(const ps_config *cfg, double mu, double var, float *table /*[R]*/);
void ps_pos_score_table(int height, int width, double mu_x, double mu_y, double var_x, double var_y,
                        double root_x, double root_y, float *table /*[H][W]*/);
void ps_torso_prior_table(int height, int width, double mu_x, double mu_y, double var_x, double var_y,
                          float weight, float *table /*[H][W]*/);

/* ---- inference -------------------------------------------------------------------------------- */

/* computeRootPosteriorRot (findrot.cpp:470-727) on the resident unaries: upright masking, border strip,
 * upward pass, downward pass (computePartMarginals, :124-286), per-part argmax, root rotation-marginal.
 * The unaries are masked in place exactly as the reference mutates its argument unless
 * PS_INFER_KEEP_UNARIES is set. Asynchronous on the ctx stream; getters synchronise. */
int ps_infer(ps_ctx *ctx, int flags);

/* getMaxStates (findrot.cpp:73-110): use_pairwise:false shortcut -- argmax (+ local maxima) of the
 * unaries at scale 0. */
int ps_max_states(ps_ctx *ctx, int flags);

/* best_part_hyp[p][0].toVect() for every part -> out[P][7] (the `best_conf` of findrot.cpp:1005-1011).
 * Describes the LAST scale, like the reference (best_part_hyp is refilled per scale, :257-259). */
int ps_get_best_conf(ps_ctx *ctx, float *out);

/* best_part_hyp[part]: the argmax record followed by <=K local maxima (findrot.cpp:277-283), rows of 7.
 * Needs PS_INFER_LOCAL_MAX.  Local maxima are ordered by the reference scan order (rotation, x, y) when
 * there are <=K of them, by descending score otherwise (ties: scan order; the reference's std::sort
 * leaves that unspecified).  *count receives the number of rows written (<= cap). */
int ps_get_part_hyps(ps_ctx *ctx, int part, float *out, int cap, int *count);

/* log_part_posterior[part] of `scale` -> dst[R][H][W] (what bSaveMarginals dumps as log_prob_grid,
 * findrot.cpp:239-253).  scale must be the last one unless keep_all_scales was set. */
int ps_get_marginal(ps_ctx *ctx, int part, int scale, float *dst, int mem_kind);

/* root_part_posterior -> dst[S][H][W] (findrot.cpp:714-726). */
int ps_get_root_posterior(ps_ctx *ctx, float *dst, int mem_kind);

/* findLocalMax(root_part_posterior, hypothesis_list, 1000) (findrot.cpp:1037-1038; aux.cpp:266-290):
 * rows of (scaleidx, x, y, score); needs PS_INFER_ROOT_HYPS. */
int ps_get_root_hyps(ps_ctx *ctx, float *out, int cap, int *count);

/* ---- test seam --------------------------------------------------------------------------------- */

/* computeRotJointMarginal (findrot.cpp:292-456) on caller buffers: one message
 * log_prob_parent = MSG(log_prob_child; offset_in, offset_out, C, rot_mean, rot_sigma, scale, sparse).
 * Upward calls pass (offset_c, offset_p, +rot_mean, sparse); downward (offset_p, offset_c, -rot_mean, 0). */
int ps_message(ps_ctx *ctx, const float *log_prob_child, float *log_prob_parent, int mem_kind,
               const double offset_in[2], const double offset_out[2], const double C[4],
               double rot_mean, double rot_sigma, double scale, int sparse);

/* computePosJointMarginal (objectdetect_findpos.cpp:64-89), the message of the legacy POS_GAUSSIAN joints, applied to
 * each of the ctx's num_rotation_steps slices of [D][H][W] independently (the reference calls it on 2-D grids):
 * exp without max shift, gaussFilter2dOffset(C * scale^2, offset * scale, unnormalised, sparse) -- always through the
 * eigen-frame, the offset on the back transform (multi_array_filter.hpp:335-369) -- and log.  Like the reference it
 * also rewrites `log_prob_child` with log(exp(child)). */
int ps_pos_message(ps_ctx *ctx, float *log_prob_child, float *log_prob_parent, int mem_kind, const double offset[2],
                   const double C[4], double scale, int sparse);

/* findLocalMax core (aux.cpp:193-261) on a caller grid [D0][H][W] (H, W may differ from the ctx's):
 * rows of (dim0, x, y, score). */
int ps_find_local_max(ps_ctx *ctx, const float *grid, int mem_kind, int d0, int height, int width,
                      int max_n, float *out, int *count);

/* What ps_set_joints derived for one message (joint, direction, scale):
 * out = {diagonal covariance?, filter-grid rows, filter-grid cols (the eigen-frame grid of gaussFilter2dOffset,
 * transform.hpp:298-299, or the image grid), rotation taps (0: no blur), x taps, y taps, rot_mean_idx, shift flags,
 * cells per rotation slice the x pass computes, cells per slice the y pass computes (the work lists keep only what
 * the read-back can reach; rows*cols when no list applies)}.
 * Used by bench.py to count tap-outputs for the fp32-pipe roofline. */
int ps_get_plan_info(ps_ctx *ctx, int joint, int downward, int scale, int out[10]);

/* Host-only (no GPU, no ctx): the work lists of the two Gaussian passes that ps_set_joints derives for a message with
 * covariance C (row-major 2x2) at `scale` on the cfg's height x width grid -- which cells of the eigen-frame grid of
 * gaussFilter2dOffset (multi_array_filter.hpp:335-369) are filtered at all.
 * dims = {EH, EW, x reach (taps - 1)/2, y reach, x-list entries, y-list entries}; T34 = rows 0,1 of the affine map from
 * image to eigen-frame coordinates that the bilinear read-back uses (filter.hpp:364-368); each list entry is
 * (first row along the filtered axis, 64-cell strip of the other axis, 8-row groups).  At most `cap` entries are
 * written to each list.  Returns PS_ERR_INVALID for a diagonal covariance (no eigen-frame, no lists).
 * Lets the CPU tests check by brute force that every cell the read-back can touch is covered. */
int ps_plan_work_lists(const ps_config *cfg, const double C[4], double scale, int dims[6], double T34[6],
                       int *xlist, int *ylist, int cap);

/* Host-only (no GPU, no ctx): the work of the fused x+y Gaussian kernel for the same message -- one walk per 64-column
 * strip of the eigen-frame grid.  dims = {EH, EW, x reach, y reach, halo (y reach rounded up to 8), lag K, walks}.
 * walks[4*i..] = (strip, first row, 8-row groups, offset of the walk's x-block masks in `masks`): the y filter runs over
 * rows [first row, first row + 8*groups) of the strip; x block j of the walk covers rows first row - halo + 64 j ..+63
 * and masks[offset + j] says which of the strip's eight 8-column groups are x-filtered there (groups + 7)/8 + K blocks
 * per walk).  At most `cap` walks / `mask_cap` mask bytes are written; *nmasks receives the number of mask bytes. */
int ps_plan_walks(const ps_config *cfg, const double C[4], double scale, int dims[7], int *walks, int cap,
                  unsigned char *masks, int mask_cap, int *nmasks);

/* Exhaustive check of the device exp/log used on the path: for every fp32 bit pattern in
 * [first_bits, first_bits + count) compares the table-driven fast evaluation with CUDA's fp64 libm narrowed to fp32
 * (the reference calls the double libm routines on floats: multi_array_op.hpp:165,177).
 * out = {exp mismatches, exp inputs tested, log mismatches, log inputs tested}. */
int ps_selftest_math(ps_ctx *ctx, unsigned first_bits, unsigned long long count, unsigned long long out[4]);

/* Evaluates the device exp/log on the fp32 bit patterns [first_bits, first_bits + count) into a host buffer:
 * op 0 = exp as used on the path, 1 = log as used on the path, 2 / 3 = CUDA's fp64 exp / log narrowed to fp32.
 * Lets a test compare them with the host libm the oracle uses. */
int ps_eval_math(ps_ctx *ctx, int op, unsigned first_bits, unsigned count, float *out_host);

/* Number of CUDA kernels this ctx has launched so far (bench.py's gpu_launches). */
long long ps_launch_count(const ps_ctx *ctx);

/* Device timing per kernel class: with profiling on, every launch on the ctx stream is bracketed by CUDA events.
 * ps_profile_read synchronises, returns (name, total ms, launches) per class seen since the last read and resets.
 * The reference has no counterpart (its only timer, get_runtime(), findrot.cpp:52-57, is commented out at :987-998). */
int ps_profile_enable(ps_ctx *ctx, int on);
int ps_profile_read(ps_ctx *ctx, int cap, const char **names, double *total_ms, long long *launches, int *count);

/* Library identification: "psinfer <version> sm_100a". */
const char *ps_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSINFER_H_ */
