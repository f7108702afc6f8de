/*
 * qcart_rollout.h -- the device-resident steps either side of the SSE hot path, exported by
 * the same libqcart.so.  They remove the host round trip the reference makes once per control step and per actor:
 *
 *   obs = get_data(state) * input_scaling  -> shared buffer -> pipe -> manager -> net(network_input) -> argmax -> pipe -> actor
 *   (quartic oscillator/main_parallel.py:128-131,150-165,210,331-360)
 *
 * Here the moment block written by qc_step stays in HBM: qc_obs_f32 turns it into the float32 observation, qc_policy_forward
 * evaluates the reference's `direct_DQN` (quartic oscillator/RL.py:81-112, layers.py:7-80,97-103) on the whole batch,
 * qc_epsilon_greedy applies the actor's exploration rule, and the int32 action tensor is consumed by the next qc_step without leaving
 * the device.  qc_replay_* writes the reference's experience rows (main_parallel.py:212-215) into a device ring; qc_record_* keeps
 * the sliding window of coarse-grained measurement outcomes of `--input measurements` (harmonic oscillator/main_parallel.py:142-150,
 * 259-292).
 *
 * Conventions are those of qcart.h: plain pointers and sizes, 0 / negative qc_status, qc_last_error(), no CPU fallback.  All data
 * pointers are DEVICE pointers unless a parameter is called `host`; all work is ordered on `stream`.
 */
#ifndef QCART_ROLLOUT_H
#define QCART_ROLLOUT_H

#include "qcart.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- observation --------------------------------------------------------------------------------------------------------
 * obs[b,k] = float32(moments[b,k]) * float32(input_scaling)     (`get_data(state)*args.input_scaling`, Q/main_parallel.py:128-131,210) */
int qc_obs_f32(const double *moments, int64_t count, double input_scaling, float *obs, void *stream);

/* ---- policy: direct_DQN ---------------------------------------------------------------------------------------------------
 * Topology (Q/RL.py:81-105):  x[n_in] -fc1-> 512 -relu- fc2-> 512 -relu-+- fc31 (noisy) -> 256 -relu- fc41 (noisy) -> n_actions  (action values)
 *                                                                        +- fc32 -> 128 -relu- fc42 -> 1                           (mean prediction)
 * Parameters are float32, row-major [out, in] like torch.  Weight-normalised layers (layers.py:97-103) are passed as their EFFECTIVE
 * weight  weight / ||weight||_F * weight_norm  (the host mirror folds it); noisy layers (layers.py:7-80) as u_w, sigma_w, u_b, sigma_b.
 * With noisy_layers = 1 fc31 is a plain layer (SW/SB absent), with 0 both are. */
typedef struct qc_policy qc_policy;

enum {
    QC_P_FC1_W = 0, QC_P_FC1_B, QC_P_FC2_W, QC_P_FC2_B,
    QC_P_FC31_UW, QC_P_FC31_SW, QC_P_FC31_UB, QC_P_FC31_SB,
    QC_P_FC41_UW, QC_P_FC41_SW, QC_P_FC41_UB, QC_P_FC41_SB,
    QC_P_FC32_W, QC_P_FC32_B, QC_P_FC42_W, QC_P_FC42_B,
    QC_P_COUNT
};

/* how the factorised noise of the noisy layers is obtained */
enum {
    QC_NOISE_OFF = 0,       /* `noisy = False`: F.linear(x, u_w, u_b)  (layers.py:38-40) */
    QC_NOISE_GIVEN = 1,     /* per-sample rows [B, qc_policy_noise_width()] = (rand_in31[512], rand_out31[256], rand_in41[256], rand_out41[n_actions],
                               pad to a multiple of 4 floats; the base pointer must be 16-byte aligned),
                               i.e. the entries of the reference's randbuffer_in/out (already passed through f(x) = sign(x) sqrt|x|, layers.py:33-35,82-83) */
    QC_NOISE_PHILOX = 2     /* drawn in a kernel: Philox4x32-10 normals keyed by (seed; traj_offset + b, counter), then f(x) */
};

int qc_policy_create(int32_t n_in, int32_t n_actions, int32_t noisy_layers, int32_t device, qc_policy **out);
int qc_policy_destroy(qc_policy *p);
int64_t qc_policy_param_size(const qc_policy *p, int32_t which);      /* floats; 0 if the tensor does not exist for this topology; <0 on error */
int qc_policy_set_param(qc_policy *p, int32_t which, const float *host, int64_t count);
int64_t qc_policy_noise_width(const qc_policy *p);

/* One batched forward pass (Q/main_parallel.py:357-358: `action_values, avg_value, _ = net(network_input); actions = action_values.max(1)[1]`).
 * obs [B, n_in] float32.  Outputs (each may be NULL): q [B, n_actions], value [B] (mean-prediction head), greedy [B] int32 argmax.
 * Arithmetic: fp32 like the reference's torch modules -- the two 512-wide layers as 3xTF32 products on the tcgen05 tensor cores with fp32
 * accumulation (error ~2^-21 per product), the input and output layers as fp32 FMA. */
int qc_policy_forward(qc_policy *p, const float *obs, int64_t B, int32_t noise_mode, const float *noise, uint64_t seed,
                      int64_t traj_offset, uint64_t counter, float *q, float *value, int32_t *greedy, void *stream);

/* The actor's epsilon-greedy rule (Q/main_parallel.py:150-165): with probability eps a uniform action in [0, n_actions), else the greedy one.
 * Draws are Philox uniforms keyed by (seed; traj_offset + b, counter).  random_flag (nullable) receives the reference's `rnd`. */
int qc_epsilon_greedy(const int32_t *greedy, int64_t B, int32_t n_actions, double eps, uint64_t seed, int64_t traj_offset,
                      uint64_t counter, int32_t *action, uint8_t *random_flag, int32_t device, void *stream);

/* force[b] = (action[b] - (n_levels-1)/2) * f_max / ((n_levels-1)/2)    (`convert_to_force`, Q/RL.py:108-112) */
int qc_action_forces(const int32_t *action, int64_t B, int32_t n_levels, double f_max, double *force, int32_t device, void *stream);

int64_t qc_policy_launch_count(const qc_policy *p);
/* Hidden-layer GEMM kernel: 0 (default) = tcgen05 tensor cores with a 3xTF32 split (fp32-level accuracy, fp32 accumulation in tensor memory),
 * operands pre-split in global memory and fed by TMA; 1 = fp32 FMA on the CUDA cores (the cross-check; bit-for-bit fp32 products);
 * 2 = the tensor-core arithmetic of 0 with operands split and staged by the CTA's threads; 3 = TMA-fed from the plain fp32 matrices with
 * the low parts derived in shared memory.  0, 2 and 3 give bitwise identical results. */
int qc_policy_set_gemm(qc_policy *p, int32_t kind);

/* ---- experience rows --------------------------------------------------------------------------------------------------------
 * A ring of float32 rows [capacity, row_len].  qc_replay_push appends, for every trajectory with keep[b] != 0 and in trajectory order,
 *   row = ( last_obs[b, 0:K], obs[b, 0:K], float(last_action[b]), float(reward_scale * reward_src[b * reward_stride]) )
 * which is the reference's `np.hstack((last_data, data, [last_action], [-energy*reward_multiply]))` (Q/main_parallel.py:212-215) with
 * row_len = 2K + 2.  The write cursor lives on the device; nothing synchronises with the host. */
typedef struct qc_replay qc_replay;
int qc_replay_create(int32_t row_len, int64_t capacity, int32_t device, qc_replay **out);
int qc_replay_destroy(qc_replay *r);
int qc_replay_push(qc_replay *r, const float *last_obs, const float *obs, int32_t K, const int32_t *last_action, const double *reward_src,
                   int64_t reward_stride, double reward_scale, const uint8_t *keep, int64_t B, void *stream);
int qc_replay_total(qc_replay *r, int64_t *total_rows_pushed, void *stream);      /* synchronises `stream` */
float *qc_replay_data(qc_replay *r);                                              /* device pointer of the ring */
int qc_replay_read(qc_replay *r, int64_t first_row, int64_t n_rows, float *host, void *stream);   /* rows by absolute index (mod capacity); synchronises */

/* ---- measurement record (`--input measurements`) -----------------------------------------------------------------------------------
 * Per trajectory the last read_length + control_len coarse-grained outcomes  mean(q over coarse_grain substeps) * input_scaling  and the
 * last read_length/control_len + 1 applied forces * input_scaling, both float32 (H/main_parallel.py:259-292).  control_len is the
 * reference's read_control_step_length = control_interval / coarse_grain.  New trajectories start from all zeros (:260-261). */
typedef struct qc_record qc_record;
int qc_record_create(int64_t B, int32_t read_length, int32_t coarse_grain, int32_t control_len, int32_t device, qc_record **out);
int qc_record_destroy(qc_record *r);
int qc_record_reset(qc_record *r, const uint8_t *mask, void *stream);            /* mask NULL = all trajectories */
/* Append one control interval: q [B, n_sub] as written by qc_step (n_sub = control_len * coarse_grain), force [B] applied during it. */
int qc_record_push(qc_record *r, const double *q, int32_t n_sub, const double *force, double input_scaling, void *stream);
/* Network input [B, 2, read_length], newest first: row 0 the outcomes after discarding the oldest control_len, row 1 the force that was
 * applied while each was taken (H/main_parallel.py:279-283). */
int qc_record_window(qc_record *r, float *out, void *stream);
/* Measurement part of an experience row [B, read_length + control_len + read_length/control_len + 1], newest first (H/main_parallel.py:271-274). */
int qc_record_experience(qc_record *r, float *out, void *stream);
int64_t qc_record_row_len(const qc_record *r);

#ifdef __cplusplus
}
#endif
#endif
