/*
 * protonoc_b200.h -- C ABI of the B200-native crime / recruitment hot path of PyPROTON-OC.
 *
 * The reference (LABSS/PyPROTON-OC) is pure Python and has no FFI; its boundary for this path is
 * the ProtonOC / Person method set.  Each entry point below names the reference method(s) it
 * replaces (paths relative to the reference root).  The library owns only device state inside a
 * poc_ctx; every pointer argument is caller-owned HOST memory unless stated otherwise (pinned host
 * memory makes the copies asynchronous).  All functions return 0 on success or a negative
 * poc_status; the message is available from poc_last_error().  Nothing throws across the ABI.
 * A ctx is single-threaded and owns one CUDA stream; calls are asynchronous until poc_sync(), a
 * download, or a call with a host output argument.
 *
 * One ctx holds a BATCH of `n_replicas` independent simulations (the override-mode ensemble,
 * protonoc/run_modes.py:221-232): every kernel processes all replicas of the batch in one launch.
 * A single simulation is a batch of one.
 *
 * Slots: agents are addressed by dense slot index 0..n-1 in ascending unique_id order (the Mesa
 * schedule order the reference iterates in, model.py:1430).  Agents that died but are still
 * referenced from someone's neighbour set stay as slots with alive = 0.
 */
#ifndef PROTONOC_B200_H
#define PROTONOC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POC_NUM_LAYERS 9 /* Person.network_names order, entities.py:38-47 */
enum poc_layer {
    POC_SIBLING = 0, POC_OFFSPRING = 1, POC_PARENT = 2, POC_PARTNER = 3, POC_HOUSEHOLD = 4,
    POC_FRIENDSHIP = 5, POC_CRIMINAL = 6, POC_PROFESSIONAL = 7, POC_SCHOOL = 8
};
#define POC_MAX_CELLS 32
#define POC_MAX_TAB 16
#define POC_MAX_GROUP 32

enum poc_status {
    POC_OK = 0, POC_ERR_CUDA = -1, POC_ERR_ARG = -2, POC_ERR_CAPACITY = -3, POC_ERR_STATE = -4, POC_ERR_NOGPU = -5
};

typedef struct poc_ctx poc_ctx;

/* Model parameters and tables read by the hot path (model.py:108-134, 159-190).  CDFs are built on the
 * host exactly as extra.weighted_n_of + numpy Generator.choice build them (extra.py:114-128). */
typedef struct poc_params {
    int32_t n_cells;                       /* crime_rate_by_gender_and_age_range.csv rows, file order */
    int32_t cell_male[POC_MAX_CELLS];
    int32_t cell_age_from[POC_MAX_CELLS];
    int32_t cell_age_to[POC_MAX_CELLS];
    double cell_c[POC_MAX_CELLS];
    int32_t n_sizes;                       /* num_co_offenders_dist: group size n and its CDF */
    int32_t size_n[POC_MAX_TAB];
    double size_cdf[POC_MAX_TAB];
    int32_t n_sent;                        /* conviction_length: months and per-sex CDF */
    double sent_months[POC_MAX_TAB];
    double sent_cdf_m[POC_MAX_TAB];
    double sent_cdf_f[POC_MAX_TAB];
    double propensity_threshold;           /* entities.py:350-353 evaluated on the host */
    double crimes_yearly_per10k;
    double arrests_per_year;
    double punishment_length;
    double threshold_use_facilitators;
    double facilitator_repression_multiplier;
    double good_guy_threshold;
    int32_t max_accomplice_radius;
    int32_t oc_embeddedness_radius;
    int32_t facilitator_repression;
    int32_t oc_boss_repression;
    int32_t ticks_between_intervention;
    int32_t intervention_start;
    int32_t intervention_end;
    int32_t this_is_a_big_crime;
    int32_t ticks_per_year;
    int32_t _pad;
} poc_params;

/* Structure-of-arrays view of the Person attributes the path reads or writes (entities.py:49-83). */
typedef struct poc_agent_cols {
    uint8_t *alive, *male, *oc_member, *facilitator, *prisoner, *has_job, *has_school, *target;
    int8_t *job_level, *education_level, *wealth_level;
    int32_t *age, *birth_tick, *num_crimes, *num_crimes_tick, *father, *new_recruit;
    double *propensity, *tendency, *sentence, *arrest_weight;
} poc_agent_cols;

/* The reference's recorded random draws for one tick of one replica.  Used for the
 * bit-exact parity mode; production draws come from Philox4x32-10 keyed by (philox_key; tick, stage, index). */
typedef struct poc_replay {
    int32_t n_crimes;
    const double *u_init;    /* initiator draw per crime, extra.py:135-137 */
    const double *u_size;    /* group-size draw per crime, model.py:1507-1513 */
    const int32_t *fac_pick; /* chosen facilitator slot per crime or -1, entities.py:444-445 */
    double u_arrest_count;   /* model.py:1481 */
    int32_t n_u_arrest;
    const double *u_arrest;  /* uniforms of the no-replacement sampler, extra.py:132 */
    int32_t n_u_sentence;
    const double *u_sentence;/* one per arrest, entities.py:590-593 */
    int32_t n_arrest_idx;    /* >= 0: use these indices into the flattened criminals list verbatim */
    const int32_t *arrest_idx;
} poc_replay;

typedef struct poc_tick_out {
    int32_t n_crimes, n_members, n_arrests, n_emb_evals;
    int32_t d_number_crimes, d_crime_size_fails, d_facilitator_crimes, d_facilitator_fails;
    int32_t d_big_crime_from_small_fish, d_offspring_recruited, d_protected_recruited, d_people_jailed;
    int32_t n_recruited, error;
    int64_t traversed_edges;
} poc_tick_out;

/* Optional per-crime records of one tick (host buffers sized by the caller). */
typedef struct poc_tick_records {
    int32_t cap_crimes;       /* capacity of initiator/k/group_off(+1) */
    int32_t cap_members;      /* capacity of group_mem */
    int32_t cap_arrests;
    int32_t *initiator;       /* slot per crime */
    int32_t *k;               /* accomplices requested per crime */
    int32_t *group_off;       /* CSR over group_mem; members ascending slot, initiator last */
    int32_t *group_mem;
    int32_t *arrested;        /* slots in arrest order */
} poc_tick_records;

/* ---- lifecycle ------------------------------------------------------------------------------- */
int poc_device_count(void);
int poc_ctx_create(int device, int n_replicas, int max_agents, int64_t max_static_entries,
                   int64_t max_criminal_entries, poc_ctx **out);
int poc_ctx_destroy(poc_ctx *ctx);
const char *poc_last_error(poc_ctx *ctx);
int poc_sync(poc_ctx *ctx);
/* CUDA-event time (ms) spent in kernels of the named class since the last reset; see bench.py */
int poc_set_timing(poc_ctx *ctx, int on);
int poc_timers_reset(poc_ctx *ctx);
int poc_timers_read(poc_ctx *ctx, double *ms_out /* [8] */, int64_t *launches_out /* [8] */);
/* kernel classes of the timers: 0 cells, 1 tendency, 2 draw, 3 search, 4 embeddedness, 5 commit, 6 arrests, 7 other */
long long poc_launch_count(poc_ctx *ctx); /* kernels launched since the last poc_timers_reset */
/* CUDA events on the ctx stream for callers that time a region on the device (16 slots) */
int poc_marker(poc_ctx *ctx, int slot);
int poc_marker_elapsed_ms(poc_ctx *ctx, int from_slot, int to_slot, double *ms_out);
/* raw per-replica counters [n_replicas][24] and traversed multiplex entries [n_replicas] */
int poc_read_counters(poc_ctx *ctx, int32_t *counters_out, uint64_t *edges_out);
/* algorithmic-work counters [n_replicas][16]: embeddedness requests, full evaluations, algorithmic bytes, ball nodes;
 * accomplice searches, algorithmic bytes, candidates scored; spare; 8..15 per-phase cycles of -DPOC_PHASES builds.  reset != 0 zeroes them (and the edge counter). */
int poc_read_stats(poc_ctx *ctx, uint64_t *stats_out, int reset);

/* ---- state in / out (replaces the Person object graph, entities.py:36-108) -------------------- */
int poc_set_params(poc_ctx *ctx, int replica /* -1: all */, const poc_params *p);
int poc_upload_agents(poc_ctx *ctx, int replica, int n, const poc_agent_cols *cols);
/* rows sorted ascending by col; co_off only for POC_CRIMINAL (Person.num_co_offenses, entities.py:83).
 * poc_upload_agents / poc_upload_layer are asynchronous: the caller's buffers must stay valid, and invalid input is
 * reported, when poc_build_graph (which synchronises) returns. */
int poc_upload_layer(poc_ctx *ctx, int replica, int layer, const int32_t *rowptr, const int32_t *col,
                     const int32_t *co_off_or_null);
/* Packed state: one contiguous host block per state (header, the agent columns, nine {rowptr, col} pairs, co-offence
 * counts) built once with poc_pack_state and uploaded with ONE copy per replica by poc_upload_packed -- how an
 * ensemble driver (run_modes.py:221-232) hands hundreds of initial states to one ctx.  poc_packed_size: bytes needed
 * for n agents with n_entries[l] directed entries in layer l.  poc_upload_packed is asynchronous (the block must stay
 * valid until poc_sync, which also reports what the device-side input checks found); pinned memory makes the copy
 * overlap the unpacking of the previous replica. */
int64_t poc_packed_size(int n, const int64_t *n_entries /* [POC_NUM_LAYERS] */);
int poc_pack_state(int n, const poc_agent_cols *cols, const int32_t *const *rowptr, const int32_t *const *col,
                   const int32_t *co_off_or_null, void *block, int64_t block_bytes);
int poc_upload_packed(poc_ctx *ctx, int replica, const void *block, int64_t block_bytes);
/* merge the eight non-criminal layers into the packed multiplex CSR the kernels traverse */
int poc_build_graph(poc_ctx *ctx, int replica);
/* the same without waiting: the uploads of the next replica queue up behind this one.  Caller-owned page-locked
 * buffers must stay valid until poc_sync(), which also reports invalid input found by the deferred checks. */
int poc_build_graph_async(poc_ctx *ctx, int replica);
int poc_download_agents(poc_ctx *ctx, int replica, poc_agent_cols *cols);
/* the same for `count` consecutive replicas in one call (cols[i] receives replica replica0 + i): the unpack kernel and the
 * device-to-host copy of the next replicas run while the columns of the one before are handed out -- what an ensemble
 * driver calls after a run (run_modes.py:73-88 collects every run's outputs) */
int poc_download_agents_many(poc_ctx *ctx, int replica0, int count, const poc_agent_cols *cols);
int poc_layer_entries(poc_ctx *ctx, int replica, int layer, int64_t *out);
int poc_download_layer(poc_ctx *ctx, int replica, int layer, int32_t *rowptr, int32_t *col, int32_t *co_off_or_null);

/* ---- the hot path (all replicas of the batch per call) ---------------------------------------- */
/* model.py:236-238: zero this-tick counters, age = floor((tick - birth_tick)/12) for scheduled agents */
int poc_begin_tick(poc_ctx *ctx, int tick);
/* model.py:279-283: prisoner countdown */
int poc_end_tick(poc_ctx *ctx);
/* ProtonOC.calculate_criminal_tendency, model.py:1160-1186 + Person.update_criminal_tendency */
int poc_tendency(poc_ctx *ctx);
/* ProtonOC.calculate_crime_multiplier, model.py:1144-1157; result kept on device, copied to out if non-NULL */
int poc_crime_multiplier(poc_ctx *ctx, double *out_or_null /* [n_replicas] */);
int poc_set_crime_multiplier(poc_ctx *ctx, int replica, double value);
/* ProtonOC.reset_oc_embeddedness + commit_crimes, model.py:709-718, 1417-1504.
 * replay: NULL, or n_replicas pointers (each NULL => Philox for that replica). */
int poc_commit_crimes(poc_ctx *ctx, int tick, const poc_replay *const *replay_or_null, const uint64_t *philox_keys,
                      poc_tick_out *out_or_null /* [n_replicas] */, poc_tick_records *records_or_null /* replica 0 */);
/* n_ticks iterations of {begin_tick; yearly tendency + multiplier; commit_crimes; end_tick}
 * (the hot-path subset of ProtonOC.step, model.py:226-288) with no host synchronisation inside.
 * reporters_out: [n_replicas][n_ticks][POC_N_REPORTERS] doubles, copied back at the end (may be NULL). */
#define POC_N_REPORTERS 43 /* the numeric columns of calculate_fast_reporters + the path's counters, see engine.REPORTER_NAMES */
int poc_run_ticks(poc_ctx *ctx, int tick0, int n_ticks, const uint64_t *philox_keys, double *reporters_out);
/* the same, but the reporter rows stay on the device (for a collective over NVLink, run_modes.py:221-232 gathers the
 * outputs of the runs): *rows_dev points at [n_replicas][*ticks_per_replica][POC_N_REPORTERS] doubles owned by the ctx and
 * valid until the next run; rows 0..n_ticks-1 of every replica are this run's.  Returns after the run has completed. */
int poc_run_ticks_device(poc_ctx *ctx, int tick0, int n_ticks, const uint64_t *philox_keys, double **rows_dev,
                         int *ticks_per_replica);
/* ProtonOC.calculate_fast_reporters, model.py:1687-1735, for the current device state: one row of POC_N_REPORTERS
 * doubles per replica (columns: engine.REPORTER_NAMES). */
int poc_report(poc_ctx *ctx, double *out /* [n_replicas][POC_N_REPORTERS] */);

/* ---- stage hooks: the Person methods the reference's tests call directly ----------------------- */
/* Person.agents_at_distance (exact = 1) / agents_in_radius (exact = 0), entities.py:482-524.
 * out_off[n_src + 1], members ascending slot. */
int poc_rings(poc_ctx *ctx, int replica, int n_src, const int32_t *src, int d, int exact, int32_t *out_off,
              int32_t *out_mem, int64_t cap_mem);
/* Person.social_proximity, entities.py:674-689 */
int poc_social_proximity(poc_ctx *ctx, int replica, int n_pairs, const int32_t *a, const int32_t *b, double *out);
/* Person.oc_embeddedness, entities.py:526-558 (own-ball meta graph) */
int poc_embeddedness(poc_ctx *ctx, int replica, int n, const int32_t *agents, double *out);
/* Person.oc_embeddedness exactly as the reference computes it inside one commit_crimes (test hook): the
 * evaluations `agents`, IN THIS ORDER, on the tick-global meta graph that every evaluation adds its ball to
 * (model.py:709-718, 1515-1536; entities.py:545-558).  over: [n_over][4] rows (evaluation index, a, b, multiplicity):
 * after that evaluation the meta edge {a,b} carries 1/multiplicity (the direction nx.Graph.add_edge saw last on a pair
 * whose two directed multiplicities differ); NULL / 0 = none.  The values also become the per-tick cache
 * (entities.py:531) of the NEXT poc_commit_crimes on this ctx, so a replayed tick ranks candidates with them. */
int poc_embeddedness_accumulated(poc_ctx *ctx, int replica, int n, const int32_t *agents, int n_over, const int32_t *over,
                                 double *out);
/* test hook: hand the next poc_commit_crimes its per-tick oc_embeddedness cache (entities.py:531) directly */
int poc_preload_embeddedness(poc_ctx *ctx, int replica, int n, const int32_t *agents, const double *values);
/* Person.candidates_weight, entities.py:457-467 */
int poc_candidate_weights(poc_ctx *ctx, int replica, int n_pairs, const int32_t *self, const int32_t *cand, double *out);
/* Person.find_accomplices, entities.py:413-455: n independent searches on the current state.
 * groups: [n][POC_MAX_GROUP] (selection order, initiator last), group_len[n], d_fails[n][3]. */
int poc_find_accomplices(poc_ctx *ctx, int replica, int n, const int32_t *initiator, const int32_t *k,
                         const int32_t *fac_pick_or_null, const double *u_fac_or_null, int32_t *groups,
                         int32_t *group_len, int32_t *d_fails);

/* ---- the phases of ProtonOC.step that write the multiplex network, on the device ------------------
 * make_baby (model.py:1553-1573, entities.py:605-613, 628-650), make_friends (model.py:610-626, entities.py:246-254),
 * make_people_die (model.py:1575-1602, entities.py:615-626, 652-661) with a batch semantics (every agent of a phase
 * decides on the state the phase started from).  Off by default: poc_run_ticks is the crime path. */
typedef struct poc_demography {
    double mort[2][128];      /* yearly death probability by [male][age] (initial_mortality_rates.csv); past the table 1 */
    double fert[3][64];       /* yearly fertility by [min(number_of_children, 2)][age] (initial_fertility_rates.csv) */
    double poisson_cdf[32];   /* CDF of Poisson(3): new friends an agent asks for (model.py:619) */
    double prop_q[1024];      /* quantiles of lognormal(nat_propensity_m, nat_propensity_sigma): a newborn's propensity */
    int32_t mort_max_age;     /* last age of the mortality table */
    int32_t n_prop;           /* quantiles in use (<= 1024) */
    int32_t retirement_age;   /* model.py:121 */
    int32_t _pad;
    double weddings_mean;     /* number_weddings_mean: marriages per 1,000 persons and year (model.py:184-186) */
} poc_demography;
/* make_baby, make_friends, make_people_die; then retire_persons (model.py:1538-1551), remove_excess_friends (:629-643),
 * remove_excess_professional_links (:646-659) and wedding (:368-402) */
enum { POC_PHASE_BABY = 1, POC_PHASE_FRIENDS = 2, POC_PHASE_DIE = 4, POC_PHASE_RETIRE = 8, POC_PHASE_XFRIENDS = 16,
       POC_PHASE_XPROF = 32, POC_PHASE_WEDDING = 64, POC_PHASE_ALL = 127 /* the seven phases of every tick */,
       POC_PHASE_LABOUR = 128 /* the yearly labour-status block, below */, POC_PHASE_EVERY = 255 };
int poc_set_demography(poc_ctx *ctx, const poc_demography *d);
/* which of the phases poc_run_ticks runs every tick, in the order of model.py:248-284 (default 0).
 * max_agents / max_static_entries of the ctx must leave room for the newborns and their links. */
int poc_set_phases(poc_ctx *ctx, int phase_mask);
/* one phase for all replicas on the current state (stage hook for the parity tests) */
int poc_run_phase(poc_ctx *ctx, int phase, int tick, const uint64_t *philox_keys);
/* Person.number_of_children (entities.py:75): defaults to the number of offspring links at upload */
int poc_upload_children(poc_ctx *ctx, int replica, const int32_t *n_children);
int poc_download_children(poc_ctx *ctx, int replica, int32_t *n_children);
/* ---- the yearly labour-status block of ProtonOC.step (model.py:248-256; SURVEY.md 8(f) "then" rows) ------------------
 * POC_PHASE_LABOUR = graduate_and_enter_jobmarket (model.py:1351-1380 with Person.enroll_to_school / leave_school,
 * entities.py:292-308, 389-395) and the update_unemployment_status loop written out in step() (model.py:252-256,
 * entities.py:397-411), on the ticks where the reference runs them (tick % ticks_per_year == 0 or tick == 1), after the
 * tendencies and the crime multiplier of that tick.  Both change the agent's own education / job / wealth level and
 * school membership only (enroll_to_school adds no links and the identity of the school matters to nothing on this
 * path), so a thread per agent with its own Philox draws is the reference's loop up to the random stream. */
typedef struct poc_labour {
    int32_t n_levels;            /* education levels = rows of schools.csv (model.py:953-965) */
    int32_t primary_age;         /* education_levels[1][0] */
    int32_t end_age[8];          /* [level] education_levels[level][1], level 1..n_levels */
    int32_t work_n[8][2];        /* work_status_by_edu_lvl[education level][male] (model.py:166-167): rows, */
    int32_t work_val[8][2][8];   /*   work status of each row, */
    double work_cdf[8][2][8];    /*   choice CDF of the rates (extra.pick_from_pair_list, extra.py:140-149) */
    int32_t wealth_n[8][2];      /* wealth_quintile_by_work_status[work status][male] (model.py:168-169) */
    int32_t wealth_val[8][2][8];
    double wealth_cdf[8][2][8];
    double inactive[2][128];     /* labour_status_by_age_and_sex[male][age] (model.py:180-181); < 0: no row */
    uint8_t range_age[2][128];   /* age is a key of labour_status_range[male] (model.py:182-183, 255) */
} poc_labour;
int poc_set_labour(poc_ctx *ctx, const poc_labour *tables);
/* Person.max_education_level (entities.py:63) and the diploma level of Person.my_school (0 = none; read only while the
 * agent's has_school is set) for every agent slot of a replica; both default to 0.  Newborns inherit the father's (else
 * the mother's) max_education_level on the device (entities.py:646-649). */
int poc_upload_education(poc_ctx *ctx, int replica, const int8_t *max_education_level, const int8_t *school_level);
int poc_download_education(poc_ctx *ctx, int replica, int8_t *max_education_level, int8_t *school_level);
int poc_sizeof_labour(void);
/* agent slots of a replica (grows with births) */
int poc_n_agents(poc_ctx *ctx, int replica, int32_t *out);

int poc_sizeof_params(void);
int poc_sizeof_tick_out(void);

#ifdef __cplusplus
}
#endif
#endif /* PROTONOC_B200_H */
