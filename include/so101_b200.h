/*
 * so101_b200.h — C ABI of the B200-native batched SO-ARM101 stepper.
 *
 * This is the drop-in boundary for ONE hot path of Hucheyu1/Lerobot-mujoco-sim2real:
 * SOARM101Env.reset/step (frame_skip x mujoco.mj_step on the SO101 hinge chain) and the
 * trajectory loop of SOARM101DataGenerator.generate_physics_based_data.
 *
 * The reference has no FFI of its own: it calls the third-party `mujoco` pybind module.
 * Each entry point below names the reference call site(s) it replaces:
 *
 *   so101_model_create      <- mujoco.MjModel.from_xml_path         [REF SOARM101/SOARM101_Env.py:34]
 *   so101_batch_create      <- mujoco.MjData(model)                 [REF SOARM101/SOARM101_Env.py:43]
 *   so101_batch_reset*      <- mj_resetData + qpos/qvel write + mj_forward
 *                                                                   [REF SOARM101/SOARM101_Env.py:87-102]
 *   so101_batch_forward     <- mujoco.mj_forward                    [REF SOARM101/SOARM101_Env.py:102,
 *                                                                    Koopman_MPC.py:90,126]
 *   so101_batch_step*       <- data.ctrl[:5]=u ; frame_skip x mj_step ; _get_state
 *                                                                   [REF SOARM101/SOARM101_Env.py:128-135, 69-75]
 *   so101_batch_rollout*    <- the (reset, T x step, row write) double loop
 *                                                                   [REF SOARM101/SOARM101_DataCollection.py:108-134]
 *   so101_batch_shoot       <- batched evaluation of control sequences from one shared state
 *                              (Koopman_MPC.py:197-222 closed loop, BASELINE.json config 5)
 *   so101_ik_track          <- the per-way-point dm_control qpos_from_site_pose loop of
 *                              CartesianTrajectoryGenerator.generate / _solve_ik
 *                                                                   [REF control/TrajectoryGenerator.py:81-116, 180-210]
 *
 * Conventions
 *   - plain C, no exceptions; every call returns 0 on success, <0 on error (so101_last_error()
 *     gives the text, thread-local).
 *   - device buffers are structure-of-arrays `[field][N]` (N = n_envs), element type = the batch
 *     dtype (SO101_F64 -> double, SO101_F32 -> float) unless stated; observations are always
 *     float32 `[8][N]` = [ee_x, ee_y, ee_z, q0..q4] (reference order, SOARM101_Env.py:69-75).
 *   - all device work is enqueued on the caller's stream (`stream` = cudaStream_t as void*,
 *     pass torch.cuda.current_stream().cuda_stream); no hidden synchronisation except in the
 *     *_host variants, which synchronise the stream before returning.
 *   - a batch is bound to one device; handles are not thread-safe.
 *   - there is no CPU fallback: every compute entry point fails if no CUDA device is present.
 */
#ifndef SO101_B200_H_
#define SO101_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO101_ABI_VERSION 8
#define SO101_NV       6   /* hinge dofs: 5 arm joints + gripper */
#define SO101_MAXBODY  8   /* world, fixed base, 6 links */
#define SO101_MAXTRIP 16   /* contact-tripwire boxes (<= 3 per link) */
#define SO101_NOBS     8   /* observation: ee_pos(3) + qpos[0:5] */
#define SO101_NU_ENV   5   /* controls exposed by SOARM101Env (gripper channel held at 0) */
#define SO101_ROW     13   /* dataset row: u(5) | ee_pos(3) | qpos(5) */

enum { SO101_F64 = 0, SO101_F32 = 1 };

/* error codes */
enum {
  SO101_OK = 0,
  SO101_EINVAL = -1,      /* bad argument */
  SO101_EMODEL = -2,      /* model outside the supported subset (not a hinge chain, ...) */
  SO101_ECUDA = -3,       /* CUDA runtime error */
  SO101_ENODEVICE = -4    /* no CUDA device: there is no CPU fallback */
};

/* per-env status bits (so101_batch_get_flags) */
enum {
  SO101_FLAG_BADSTATE   = 1u << 0,  /* qpos/qvel/qacc NaN or |x|>1e10 (mj_checkPos/Vel/Acc) */
  SO101_FLAG_TRIP_TABLE = 1u << 1,  /* a table contact that is NOT simulated occurred: no hull data loaded, more than
                                       SO101_MAXCON simultaneous contacts, or a contact beyond the table's footprint */
  SO101_FLAG_TRIP_SELF  = 1u << 2,  /* outside the self-collision-free joint box AND the oriented boxes of two non-adjacent links overlap */
  SO101_FLAG_LIMIT      = 1u << 3,  /* a joint-limit row was active at least once (info)   */
  SO101_FLAG_MAXITER    = 1u << 4,  /* Newton hit opt.iterations (info)                    */
  SO101_FLAG_CONTACT    = 1u << 5   /* table-plane contact rows were active at least once (info; simulated) */
};
#define SO101_MAXCON 6   /* simultaneous table contacts per env the kernels simulate (one per colliding geom) */

/* rollout / shoot option bits */
enum {
  SO101_ROLL_GRAVCOMP_HOLD = 1u << 0, /* qfrc_applied <- qfrc_bias at the start of every control
                                         step, held over the sub-steps [REF Koopman_MPC.py:119] */
  SO101_ROLL_NO_RESET      = 1u << 1, /* continue from the batch's current state               */
  SO101_ROLL_ROWS_F32      = 1u << 2  /* rows written as float32 instead of float64            */
};

/* control generators of SOARM101DataGenerator [REF SOARM101_DataCollection.py:97-103,115,132] */
enum {
  SO101_CTRL_RANDOM = 0,  /* u ~ U(-amp, amp) i.i.d. per control step (Philox4x32-10)        */
  SO101_CTRL_SIN    = 1,  /* u = a*sin(2*pi*f*t + phase), t = integer control step            */
  SO101_CTRL_CHIRP  = 2,  /* f = f_traj + (f_hi - f_lo) * t / T_total                         */
  SO101_CTRL_TENSOR = 3   /* u read from a device tensor [T+1][5][N] (batch dtype)            */
};

typedef struct So101CtrlSpec {
  int32_t  kind;          /* SO101_CTRL_*                                                     */
  int32_t  t_total;       /* chirp normaliser (reference: 200)                                */
  uint64_t seed;          /* Philox key                                                       */
  int64_t  env_offset;    /* global index of this shard's env 0 (results independent of the
                             number of GPUs)                                                  */
  double   amp;           /* random: half-range (0.5); sin/chirp: amplitude range +-amp       */
  double   freq_lo;       /* 0.0025                                                           */
  double   freq_hi;       /* 0.05                                                             */
  double   reset_lo;      /* reset box for qpos[0:5]: U(reset_lo, reset_hi) = (-0.3, 0.3)     */
  double   reset_hi;
  const void* u;          /* SO101_CTRL_TENSOR: device pointer [T+1][5][N]                    */
} So101CtrlSpec;

/*
 * Packed constant tables of one compiled MJCF scene (host -> library, once).  Built by the
 * Python MJCF compiler (mjcf.py); mirrors the subset of mjModel that mj_step reads for this
 * scene.  Index 0 of the body arrays is the world.  Joint j <-> dof j (hinges only).
 */
typedef struct So101Tables {
  int32_t abi_version;
  int32_t nbody;                 /* incl. world                                            */
  int32_t nv;                    /* == SO101_NV                                            */
  int32_t nu;                    /* == SO101_NV                                            */
  /* mjOption */
  int32_t iterations;            /* 100                                                    */
  int32_t ls_iterations;         /* 50                                                     */
  int32_t site_body;             /* body of the `gripperframe` site                        */
  int32_t ntrip;                 /* number of tripwire boxes                               */
  double  timestep;              /* 0.002                                                  */
  double  gravity[3];
  double  tolerance;             /* 1e-8                                                   */
  double  ls_tolerance;          /* 0.01                                                   */
  double  meaninertia;           /* mjModel.stat.meaninertia                               */
  /* bodies */
  int32_t body_parent[SO101_MAXBODY];
  int32_t body_jnt[SO101_MAXBODY];            /* joint index or -1                         */
  double  body_pos[SO101_MAXBODY][3];
  double  body_quat[SO101_MAXBODY][4];
  double  body_ipos[SO101_MAXBODY][3];
  double  body_iquat[SO101_MAXBODY][4];
  double  body_inertia[SO101_MAXBODY][3];     /* principal moments                         */
  double  body_mass[SO101_MAXBODY];
  /* joints / dofs */
  int32_t jnt_body[SO101_NV];
  int32_t jnt_limited[SO101_NV];
  double  jnt_pos[SO101_NV][3];
  double  jnt_axis[SO101_NV][3];
  double  jnt_range[SO101_NV][2];
  double  jnt_margin[SO101_NV];
  double  jnt_solref[SO101_NV][2];            /* limit rows                                */
  double  jnt_solimp[SO101_NV][5];
  double  jnt_stiffness[SO101_NV];
  double  qpos0[SO101_NV];
  double  qpos_spring[SO101_NV];
  double  dof_armature[SO101_NV];
  double  dof_damping[SO101_NV];
  double  dof_frictionloss[SO101_NV];
  double  dof_solref[SO101_NV][2];            /* friction rows                             */
  double  dof_solimp[SO101_NV][5];
  double  dof_invweight0[SO101_NV];
  double  dof_M0[SO101_NV];
  /* actuators (joint transmission, affine bias) */
  int32_t act_dof[SO101_NV];
  int32_t act_ctrllimited[SO101_NV];
  int32_t act_forcelimited[SO101_NV];
  int32_t nself;                 /* boxes [ntrip, ntrip + nself) of the trip_* arrays: colliding geoms that cannot reach the
                                    table (no tripwire box, no hull) and take part in the self-collision test only          */
  int32_t pad0_;
  double  act_gear[SO101_NV];
  double  act_gain[SO101_NV];                 /* gainprm[0]                                */
  double  act_bias[SO101_NV][3];              /* biasprm[0..2]                             */
  double  act_ctrlrange[SO101_NV][2];
  double  act_forcerange[SO101_NV][2];
  /* observation site */
  double  site_pos[3];
  /* keyframe 0 (`home`) — read by Koopman_MPC.py:65-71 through env.model */
  double  key_qpos[SO101_NV];
  double  key_ctrl[SO101_NV];
  /* contact tripwire: oriented boxes fixed to bodies vs the plane z = trip_plane_z, and a
     joint-space box outside which self-collision is possible */
  int32_t trip_body[SO101_MAXTRIP];
  double  trip_center[SO101_MAXTRIP][3];      /* body frame                                */
  double  trip_axes[SO101_MAXTRIP][9];        /* rows = box axes in the body frame         */
  double  trip_half[SO101_MAXTRIP][3];
  double  trip_plane_z;                       /* table top (world z)                       */
  double  trip_qbox[SO101_NV][2];             /* lo, hi                                    */
  /* orientation of the observation site in its body's frame (w, x, y, z): only the inverse
     kinematics (so101_ik_track) reads it, the stepper needs the site position alone */
  double  site_quat[4];
  /* table-plane contact (the scene enables contacts, scene_with_table_v.xml:28,31): colliding geom i
     (i < ntrip, same order as the tripwire boxes; its convex hull comes through so101_model_set_hulls) against the
     top face z = trip_plane_z of the static table box.  One condim-3 contact per geom (MuJoCo's convex-convex
     routine returns one), pyramidal cone -> 4 rows. */
  double  body_invweight0[SO101_MAXBODY][2];  /* mjModel.body_invweight0: translational, rotational              */
  double  con_friction[3];                    /* contact friction: elementwise max of the two geoms' friction    */
  double  con_solref[2];                      /* contact solref / solimp (solmix-weighted mean of the geoms')    */
  double  con_solimp[5];
  double  con_margin;                         /* includemargin = max margin - max gap                            */
  double  con_box[4];                         /* footprint of the table's top face: x_lo, x_hi, y_lo, y_hi       */
  int32_t con_enabled;                        /* 0: no contact tables (tripwire flags only)                      */
  int32_t con_condim;                         /* 3                                                               */
} So101Tables;

/* Convex hulls of the colliding geoms, one per tripwire box and in the same order (host pointers; copied). */
typedef struct So101Hulls {
  int32_t ngeom;               /* == So101Tables.ntrip                                                   */
  int32_t nvert;               /* total number of hull vertices                                          */
  int32_t nadj;                /* total number of adjacency entries                                      */
  int32_t cube_res;            /* G: resolution of the direction -> start-vertex cube map                */
  const int32_t* vert_start;   /* [ngeom + 1] first vertex of every geom                                 */
  const double*  vert;         /* [nvert][3] hull vertices in the frame of the geom's BODY               */
  const int32_t* adj_start;    /* [nvert + 1] CSR over the hull's edge graph                             */
  const int32_t* adj;          /* [nadj] neighbour vertex ids (global)                                   */
  const int32_t* cube;         /* [ngeom][6][G][G] a hull vertex (global id) near the support point of the direction
                                  whose largest component is +-x, +-y, +-z (faces 0..5) and whose other two components,
                                  divided by it, fall into cell (iu, iv) of [-1, 1]^2: the hill climb's start  */
} So101Hulls;

typedef struct So101Model So101Model;
typedef struct So101Batch So101Batch;

const char* so101_last_error(void);
int         so101_abi_version(void);
size_t      so101_tables_sizeof(void);
int         so101_device_count(void);

/* ---- model ---------------------------------------------------------------------------- */
int  so101_model_create(const So101Tables* tables, So101Model** out);
void so101_model_destroy(So101Model* model);
/* Attach the convex hulls of the colliding geoms: from then on batches created from this model SIMULATE table-plane
   contacts (exact support point of the hull by hill climbing on its edge graph where the bounding box dips below the
   table top -> contact rows in a general dense Newton path taken only by envs in contact) instead of only flagging
   them.  Without hulls SO101_FLAG_TRIP_TABLE marks such envs and their dynamics stay contact-free. */
int  so101_model_set_hulls(So101Model* model, const So101Hulls* hulls);

/* ---- batch ---------------------------------------------------------------------------- */
/* bytes of per-env state for n_envs (SoA rows: qpos[6] qvel[6] qacc_warmstart[6]
   qfrc_applied[6] time[1], element = batch dtype; followed by uint32 flags[N]). */
size_t so101_batch_state_bytes(int64_t n_envs, int dtype);
/* state_buf: caller-owned device memory of so101_batch_state_bytes() (e.g. a torch tensor),
   or NULL to let the library allocate. */
int  so101_batch_create(const So101Model* model, int64_t n_envs, int dtype, int device,
                        void* state_buf, So101Batch** out);
void so101_batch_destroy(So101Batch* batch);

/* Explicit experiment options of a batch (profiling sweeps, the kernel-family equality test).  The library never reads
   the environment: a stray variable cannot change which kernel runs.  value 0 = automatic for every option. */
enum {
  SO101_OPT_KERNEL_FAMILY = 0,  /* SO101_FAMILY_*: force the one-warp or the three-warp team kernels           */
  SO101_OPT_BLOCK         = 1,  /* threads per block of the one-warp kernels (multiple of 32, <= launch bound)  */
  SO101_OPT_HOST_CHUNKS   = 2,  /* pipeline depth of so101_batch_rollout_host, 1..12                            */
  SO101_OPT_HOST_EVEN     = 3,  /* 1: equal time chunks in so101_batch_rollout_host                             */
  SO101_OPT_SLICED        = 4,  /* time-sliced persistent rollout of the one-warp kernels (large batches whose env groups
                                   do not fill whole waves): 0 = automatic, 1 = always, 2 = never                    */
  SO101_OPT_REGROUP       = 5,  /* long rollouts of large batches: regroup the envs that touch the table into the same
                                   blocks every few control steps: 0 = automatic, 1 = always (one-warp kernels), 2 = never */
  SO101_OPT_HOST_DIRECT   = 7,  /* so101_batch_rollout_host: the kernels store the rows straight into the caller's pinned host
                                   buffer (no staging copy, no download phase): 0 = automatic (pinned buffer, < 128 MB of
                                   rows, >= 200 physics steps), 1 = whenever the buffer is device-accessible, 2 = never  */
  SO101_OPT_SELF_TEST     = 6   /* SO101_FLAG_TRIP_SELF: 0 = box-box test of the non-adjacent links for poses outside the
                                   fast-accept joint box (default), 1 = flag every pose outside the joint box (no test:
                                   cheaper on workloads that leave the box often, flags 10-100x more trajectories)     */
};
enum { SO101_FAMILY_AUTO = 0, SO101_FAMILY_ONEWARP = 1, SO101_FAMILY_TEAM = 2 };
int so101_batch_set_option(So101Batch* b, int option, int value);

/* mj_resetData, then qpos[0:nq] <- qpos0[j][N] (NULL: model qpos0), qvel likewise (NULL: 0),
   then mj_forward; obs (nullable) receives float32 [8][N]. */
int so101_batch_reset(So101Batch* b, const void* qpos0, const void* qvel0, void* obs, void* stream);
/* same with qpos[0:5] ~ U(lo,hi) from Philox(seed, env_offset+env), qpos[5]=0, qvel=0. */
int so101_batch_reset_random(So101Batch* b, uint64_t seed, int64_t env_offset, double lo, double hi,
                             void* obs, void* stream);
/* mj_forward at the current state: obs float32 [8][N] (nullable), qfrc_bias [6][N] (nullable). */
int so101_batch_forward(So101Batch* b, void* obs, void* qfrc_bias, void* stream);
/* ctrl [6][N] (row 5 = gripper; NULL row pointer semantics: pass n_ctrl=5 to hold it at 0);
   n_substeps x mj_step; obs float32 [8][N] with the reference's one-sub-step ee lag. */
int so101_batch_step(So101Batch* b, const void* ctrl, int n_ctrl, int n_substeps, void* obs, void* stream);
/* the same with flags: SO101_ROLL_GRAVCOMP_HOLD = qfrc_applied <- qfrc_bias of the state the env step starts from, held
   over its sub-steps - the gravity compensation line of the reference's control loops
   [REF Koopman_MPC.py:119 `data.qfrc_applied[:] = data.qfrc_bias[:]`] fused into the step launch */
int so101_batch_step_flags(So101Batch* b, const void* ctrl, int n_ctrl, int n_substeps, void* obs, uint32_t flags,
                           void* stream);
/* host-buffer variants: pinned or pageable host pointers, H2D + kernel + D2H on `stream`,
   stream synchronised before return.  These are what SOARM101Env.step()/reset() call.
   flags_host (nullable): receives the per-env status words [N] with the observation (one sync). */
int so101_batch_step_host(So101Batch* b, const void* ctrl_host, int n_ctrl, int n_substeps,
                          void* obs_host, uint32_t* flags_host, void* stream);
int so101_batch_reset_host(So101Batch* b, const void* qpos0_host, const void* qvel0_host,
                           void* obs_host, void* stream);

/* fused trajectory generation: rows [N][T+1][13] (float64, or float32 with ROWS_F32) =
   [u_i(5) | ee_pos(3) | qpos(5)], row 0 from reset, row i after step(u_{i-1}); observation
   values are rounded to float32 first, as the reference does. */
int so101_batch_rollout(So101Batch* b, const So101CtrlSpec* spec, int T, int frame_skip,
                        void* rows, uint32_t flags, void* stream);
/* host-buffer variant (what SOARM101DataGenerator calls): spec->u (SO101_CTRL_TENSOR) and qpos0_host
   [6][N] (nullable: random reset) are HOST pointers, rows_host receives the dataset.  Work is ordered
   after `stream`; long rollouts are cut into up to 12 time chunks whose uploads / downloads run on two
   internal streams and overlap the compute (invisible in the result); everything is synchronised before
   return.  Pinned host memory makes the copies async.  (Env. knob SO101_HOST_CHUNKS=1..12 overrides.) */
int so101_batch_rollout_host(So101Batch* b, const So101CtrlSpec* spec, const void* qpos0_host, int T,
                             int frame_skip, void* rows_host, uint32_t flags, void* stream);
/* B = n_envs control sequences from one shared state: state0 host pointer to 18 doubles
   (qpos, qvel, qacc_warmstart), U device [H][5][B] (batch dtype), X device float32 [B][H+1][8]. */
int so101_batch_shoot(So101Batch* b, const double* state0, const void* U, int H, int frame_skip,
                      void* X, uint32_t flags, void* stream);

/* state access, device pointers [6][N] in the batch dtype (any may be NULL) */
int so101_batch_get_state(So101Batch* b, void* qpos, void* qvel, void* qacc_warmstart, void* stream);
int so101_batch_set_state(So101Batch* b, const void* qpos, const void* qvel, const void* qacc_warmstart,
                          void* stream);
int so101_batch_set_qfrc_applied(So101Batch* b, const void* qfrc_applied, void* stream);
int so101_batch_get_flags(So101Batch* b, uint32_t* flags_dev, void* stream);
int so101_batch_clear_flags(So101Batch* b, void* stream);
/* solver statistics accumulated since the last call: [0]=physics steps, [1]=Newton iterations,
   [2]=line-search evaluations, [3]=steps with an active limit row (host pointer, 4 x uint64). */
int so101_batch_stats(So101Batch* b, uint64_t* stats_host, void* stream);

/* ---- the dataset gather without a collective ------------------------------------------------------------
   The only inter-GPU exchange of this path is the gather of the row shards [N/G][T+1][13] to rank 0
   (what np.save needs, [REF SOARM101/SOARM101_DataCollection.py:156]).  Rank 0 allocates the full [N][T+1][13] buffer
   with so101_shared_alloc and passes the 64-byte handle to the other ranks of the node (any host channel); they map it
   with so101_shared_open and hand `ptr + first_row_of_my_shard` to so101_batch_rollout as `rows`: the row writer of
   k_rollout then stores straight into rank 0's HBM over NVLink while the simulation runs - the transfer is fused into
   the kernel, nothing is staged, copied or concatenated afterwards.  Writers synchronise their stream, then all ranks
   meet at a host barrier before rank 0 reads.  One process per GPU; handles are valid on the same node only. */
#define SO101_IPC_HANDLE_BYTES 64
int so101_shared_alloc(int device, size_t bytes, void** ptr, unsigned char* handle /* [64] out */);
int so101_shared_open(int device, const unsigned char* handle /* [64] */, void** ptr);
int so101_shared_close(int device, void* ptr);
int so101_shared_free(int device, void* ptr);

/* measurement helper: register-resident FMA loop; returns achieved TFLOP/s (2 flop per FMA). */
int so101_fma_peak(int dtype, int device, double* tflops_out);

/* ---- control sequences scored under the reference's lifted linear (Koopman) model ------------------
   z_{t+1} = A z_t + B u_t from z0 (A [nz][nz], B [nz][nu], z0 [nz]: HOST, row-major; nz <= 64, nu <= 8)
   [REF models/KoopmanBase.py:49-57]; cost_b = sum_t q_weight |z_{t+1} - zref_t|^2 + r_weight |u_t|^2 with zref [H][nz]
   HOST (nullable: control effort only) [REF control/MPC_Controler.py:65-98, Q = 50 I, R = 0.5 I].  U is the DEVICE tensor
   [H][nu][n] that so101_batch_shoot takes (dtype = its element type), Xhat DEVICE float32 [n][H+1][nobs] = first nobs
   lifted coordinates (the predicted observation; nullable), cost DEVICE double [n] (nullable).  Async on `stream`. */
int so101_koopman_score(const double* A, const double* B, int nz, int nu, const double* z0, const double* zref,
                        double q_weight, double r_weight, const void* U, int H, int64_t n, int dtype, int device,
                        int nobs, void* Xhat, void* cost, void* stream);

/* ---- the body of the reference's MPC loop for n environments at once -----------------------------------
   Reference, per frame [REF Koopman_MPC.py:197-222, control/MPC_Controler.py:143-166]: z0 = Psi(state) = [x, encoder(x)]
   (MLP with ReLU between layers [REF models/KoopmanBase.py:12-47]); the H reference rows of the window lifted the same
   way (rows past the end of the trajectory stay zero in lifted space); the MPC problem ('mpc' over u, or the default
   'delta_mpc' [REF args.py:75] over delta_u with u_t = u_prev + sum delta_u) solved; u0 = u_opt[0] + u_prev;
   a = clip(u0, +-clip); u_prev <- u0 [REF Koopman_MPC.py:217].  The problem is an unconstrained quadratic, so
   u_opt[0] = Kz z0 + sum_t Kr_t zref_t + Ku u_prev with gains that depend on the model only (computed by the host side,
   koopman.py: mpc_gains).
     so101_koopman_create       uploads the encoder: dims[0..n_layers] = layer widths (dims[0] = x_dim <= 16, the others
                                multiples of 8, <= 64), W[l] HOST row-major [dims[l+1]][dims[l]] (torch Linear.weight), b[l] [dims[l+1]]
     so101_koopman_set_gains    HOST Kz [nu][nz], Kr [nu][H][nz], Ku [nu][nu] (zeros for 'mpc'); nz = dims[0] + dims[n_layers]
     so101_koopman_lift         Z [n][nz] DEVICE double <- X: layout 0 = rows [n][ldx] (first x_dim columns), 1 = structure of
                                arrays [x_dim][n] (the stepper's observation buffer); dtype of X: SO101_F64 / SO101_F32
     so101_koopman_feedforward  uff [n][P][nu] DEVICE double <- Xref [n][P][x_dim] DEVICE double:
                                uff[e][k] = sum_{t<H, k+1+t<P} Kr_t Psi(Xref[e][k+1+t]) - every reference row is lifted once
     so101_koopman_mpc_step     one frame: obs as for lift; uff = pointer to frame k of the first env (uff_stride = doubles
                                between envs, P * nu; NULL = no reference term); u_prev DEVICE double [nu][n] in/out; ctrl DEVICE
                                [nu][n] in ctrl_dtype = the control rows so101_batch_step takes; a_out DEVICE double [n][nu] or
                                NULL (the clipped controls as the dataset rows record them)
   All launches are asynchronous on `stream`. */
typedef struct So101Koopman So101Koopman;
int so101_koopman_create(int n_layers, const int32_t* dims, const double* const* W, const double* const* b, int device,
                         So101Koopman** out);
void so101_koopman_destroy(So101Koopman* k);
int so101_koopman_set_gains(So101Koopman* k, int H, int nu, const double* Kz, const double* Kr, const double* Ku);
int so101_koopman_lift(So101Koopman* k, const void* X, int dtype, int layout, int64_t ldx, int64_t n, double* Z, void* stream);
int so101_koopman_feedforward(So101Koopman* k, const double* Xref, int64_t n, int P, double* uff, void* stream);
int so101_koopman_mpc_step(So101Koopman* k, const void* obs, int obs_dtype, int obs_layout, int64_t ldx, const double* uff,
                           int64_t uff_stride, double* u_prev, void* ctrl, int ctrl_dtype, double* a_out, double clip,
                           int64_t n, void* stream);

/* ---- batched site-pose inverse kinematics along Cartesian way-point tracks ------------------------
   Replaces the loop of CartesianTrajectoryGenerator.generate [REF control/TrajectoryGenerator.py:180-210] around
   dm_control.utils.inverse_kinematics.qpos_from_site_pose [REF control/TrajectoryGenerator.py:96-107] for n
   independent tracks at once: track b visits way-points xyz[p][.][b], p = 0..P-1, in order; every solve starts from the
   previous way-point's joint vector (the first one from q0, or the model's qpos0 when q0 is NULL); a failed solve
   repeats the previous way-point's answer and restores the joint vector it started from, as the reference loop does.
   One solve = damped/min-norm Gauss-Newton on the site pose error: err = [target - site_xpos | quat2Vel(target_quat *
   conj(site_xquat))], err_norm = |err_pos| + rot_weight |err_rot|; success when err_norm < tol; update = (J'J + reg I)^-1
   J'err when err_norm > reg_threshold, else the minimum-norm least-squares solution of J'J x = J'err (singular values
   below DBL_EPSILON * largest dropped, numpy lstsq rcond=-1); stop (fail) when err_norm / |update| > progress_thresh;
   |update| clipped to max_update_norm; q += update on the dofs of dof_mask (mj_integratePos for hinges).
   All DEVICE buffers are double, structure-of-arrays over the tracks:
     xyz    [P][3][n]   way-points (world frame)
     quat   [4][n]      target orientation (w,x,y,z) held along the whole track, or NULL = position only
     q0     [6][n]      start joint vectors, or NULL
     q_out  [P][6][n]   joint vector stored for every way-point (all 6 dofs; the reference keeps [:num_joints])
     status [P][n]      int32: bit 0 = success, bit 1 = track aborted (its FIRST way-point could not be solved: the
                        reference raises RuntimeError there; the track stays at its start vector), bits 8.. =
                        iterations used (the reference's IKResult.steps)
     err    [P][n]      err_norm at exit (nullable)
   Async on `stream`; the model handle comes from so101_model_create. */
typedef struct So101IkParams {
  double  tol;                 /* 1e-6   [REF control/TrajectoryGenerator.py:103]                  */
  double  rot_weight;          /* 0.5    [REF :104]                                                */
  double  reg_strength;        /* 1e-2   [REF :105]                                                */
  double  reg_threshold;       /* 0.1    dm_control default regularization_threshold               */
  double  max_update_norm;     /* 2.0    dm_control default                                        */
  double  progress_thresh;     /* 20.0   dm_control default                                        */
  int32_t max_steps;           /* 100    [REF :102]                                                */
  int32_t dof_mask;            /* bit k = dof k may move; 0x1f = the five arm joints [REF :56]     */
} So101IkParams;

int so101_ik_track(const So101Model* model, const So101IkParams* params, const double* xyz, const double* quat,
                   const double* q0, int P, int64_t n, int device, double* q_out, int32_t* status, double* err,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SO101_B200_H_ */
