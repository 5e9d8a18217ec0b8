// Rollout -> batch assembly, native and fused with the shared-memory slot write (host code).
//
// Replaces, on the storage side of the learner,
//   RolloutAssembler.push / make_as_array      buffers/rollout_assembler.py:9-82, buffers/trajectory.py:20-42
//   LearnerStorage.make_batch                  agents/learner_storage.py:123-159
// which, per environment step, append a dict of eight 0/1-d tensors to a Python list, and per finished trajectory
// `torch.stack` the T dicts field by field, flatten each field through numpy and copy it into its slice of eight flat
// shared arrays.  Here one step record is copied ONCE, straight into row `len` of a per-trajectory staging block laid
// out exactly like a slot of the shared arrays (field-major: obs T*f | act T | rew T | logits T*A | log_prob T | is_fir T |
// hx T*H | cx T*H); a finished trajectory is eight memcpys into slot `num` of the (page-locked) shared arrays the learner
// DMA's from, plus the counter bump.
//
// Behaviour kept from the reference:
//   * trajectories are keyed by episode id; an ACTIVE trajectory older than `max_age` seconds (since its creation) is
//     dropped at the next push of any episode (policy-lag bound, rollout_assembler.py:52-56);
//   * a step of an unknown episode continues the FIRST-parked unfinished trajectory of a finished episode if there is one
//     (the reference `heappop`s an un-heapified list, i.e. takes element 0, :62-64) and marks the seam with is_fir = 1,
//     else starts a new trajectory;
//   * a trajectory of seq_len steps is emitted; an unfinished one whose episode is done is parked;
//   * a finished trajectory is written only if the shared batch has room (`num < capacity`), else it is dropped
//     (learner_storage.py:139); the counter is the reader's `batch_index`.
#include <stdint.h>
#include <string.h>
#include <deque>
#include <unordered_map>
#include <vector>
#include "common.cuh"

namespace b200rl {

struct Traj {
    double gen_time;
    int len;
    std::vector<float> data;        // one slot: field-major
};

struct Assembler {
    int T, f, A, H;
    double max_age;
    size_t off[9];                  // field offsets inside a slot (floats), off[8] = slot size
    int width[8];
    std::vector<std::pair<long long, Traj>> active;      // insertion-ordered like the reference's dict
    std::deque<std::pair<long long, Traj>> parked;       // finished episodes with unfinished trajectories, insertion order
    float* sh[8];
    int* count;
    int capacity;
    long long emitted, dropped_full, dropped_age, spliced;
};

static int find_active(Assembler* a, long long id) {
    for (size_t i = 0; i < a->active.size(); ++i) if (a->active[i].first == id) return (int)i;
    return -1;
}

}  // namespace b200rl

using namespace b200rl;

extern "C" void* b200rl_asm_create(int seq_len, int obs_width, int action_space, int hidden, double max_age_s) {
    if (seq_len < 1 || obs_width < 1 || action_space < 1 || hidden < 1) { set_error("asm_create: bad shape"); return nullptr; }
    Assembler* a = new Assembler();
    a->T = seq_len; a->f = obs_width; a->A = action_space; a->H = hidden; a->max_age = max_age_s;
    const int w[8] = {obs_width, 1, 1, action_space, 1, 1, hidden, hidden};
    size_t o = 0;
    for (int k = 0; k < 8; ++k) { a->width[k] = w[k]; a->off[k] = o; o += (size_t)seq_len * w[k]; }
    a->off[8] = o;
    for (int k = 0; k < 8; ++k) a->sh[k] = nullptr;
    a->count = nullptr; a->capacity = 0;
    a->emitted = a->dropped_full = a->dropped_age = a->spliced = 0;
    return a;
}

extern "C" int b200rl_asm_destroy(void* h) { delete (Assembler*)h; return 0; }

// the eight flat shared arrays (obs, act, rew, logits, log_prob, is_fir, hx, cx: `capacity` slots each), and the slot counter
extern "C" int b200rl_asm_bind(void* h, float* obs, float* act, float* rew, float* logits, float* log_prob, float* is_fir,
                               float* hx, float* cx, int* count, int capacity) {
    Assembler* a = (Assembler*)h;
    B200RL_REQUIRE(a && count && capacity >= 1, "asm_bind: bad arguments");
    float* p[8] = {obs, act, rew, logits, log_prob, is_fir, hx, cx};
    for (int k = 0; k < 8; ++k) { B200RL_REQUIRE(p[k], "asm_bind: null field array %d", k); a->sh[k] = p[k]; }
    a->count = count; a->capacity = capacity;
    return 0;
}

// One environment step of episode `id` at time `now_s` (monotonic seconds).  Returns 1 if this step completed a trajectory
// that was written to a slot, 2 if it completed one that had to be dropped (shared batch full), 0 otherwise, < 0 on error.
extern "C" int b200rl_asm_push(void* h, long long id, int done, double now_s, const float* obs, float act, float rew,
                               const float* logits, float log_prob, float is_fir, const float* hx, const float* cx) {
    Assembler* a = (Assembler*)h;
    if (!a || !a->count) { set_error("asm_push: assembler not bound"); return -1; }
    // expire active trajectories by age (creation time), as the reference does before every push
    for (size_t i = 0; i < a->active.size();) {
        if (now_s - a->active[i].second.gen_time < a->max_age) { ++i; continue; }
        a->active.erase(a->active.begin() + i);
        ++a->dropped_age;
    }
    int idx = find_active(a, id);
    if (idx < 0) {
        if (!a->parked.empty()) {                        // continue the first-parked unfinished trajectory, mark the seam
            a->active.emplace_back(id, std::move(a->parked.front().second));
            a->parked.pop_front();
            is_fir = 1.0f;
            ++a->spliced;
        } else {
            Traj t;
            t.gen_time = now_s; t.len = 0; t.data.assign(a->off[8], 0.f);
            a->active.emplace_back(id, std::move(t));
        }
        idx = (int)a->active.size() - 1;
    }
    Traj& t = a->active[idx].second;
    const int r = t.len;
    float* d = t.data.data();
    memcpy(d + a->off[0] + (size_t)r * a->f, obs, sizeof(float) * a->f);
    d[a->off[1] + r] = act;
    d[a->off[2] + r] = rew;
    memcpy(d + a->off[3] + (size_t)r * a->A, logits, sizeof(float) * a->A);
    d[a->off[4] + r] = log_prob;
    d[a->off[5] + r] = is_fir;
    memcpy(d + a->off[6] + (size_t)r * a->H, hx, sizeof(float) * a->H);
    memcpy(d + a->off[7] + (size_t)r * a->H, cx, sizeof(float) * a->H);
    t.len = r + 1;
    int rc = 0;
    if (t.len >= a->T) {
        const int num = *a->count;
        if (num < a->capacity) {
            for (int k = 0; k < 8; ++k) {
                const size_t n = (size_t)a->T * a->width[k];
                memcpy(a->sh[k] + (size_t)num * n, d + a->off[k], sizeof(float) * n);
            }
            __atomic_store_n(a->count, num + 1, __ATOMIC_RELEASE);      // the reader sees a complete slot
            ++a->emitted;
            rc = 1;
        } else {
            ++a->dropped_full;
            rc = 2;
        }
        a->active.erase(a->active.begin() + idx);
    } else if (done) {
        a->parked.emplace_back(id, std::move(t));
        a->active.erase(a->active.begin() + idx);
    }
    return rc;
}

// stats[6] = {active trajectories, parked, emitted, dropped (batch full), dropped (age), spliced seams}
extern "C" int b200rl_asm_stats(void* h, long long* stats) {
    Assembler* a = (Assembler*)h;
    B200RL_REQUIRE(a && stats, "asm_stats: bad arguments");
    stats[0] = (long long)a->active.size(); stats[1] = (long long)a->parked.size(); stats[2] = a->emitted;
    stats[3] = a->dropped_full; stats[4] = a->dropped_age; stats[5] = a->spliced;
    return 0;
}
