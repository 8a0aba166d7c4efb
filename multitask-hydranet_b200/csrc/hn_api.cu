// C-ABI plumbing: error string, version, device query, and the plan (a recorded schedule of ops).
#include <stdarg.h>

#include <vector>

#include "hn_common.cuh"
#include "hn_ops.h"

static thread_local char g_err[1024] = "";

void hn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int g_hn_pdl = 0;  // measured on B200 (batch 32, graph replay): 10.37 ms/step with PDL edges vs 10.06 without -> off by default
extern "C" void hn_set_pdl(int on) { g_hn_pdl = on ? 1 : 0; }

extern "C" const char* hn_last_error(void) { return g_err; }
extern "C" int hn_version(void) { return 100; }
extern "C" int hn_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
enum OpKind { OP_CONV, OP_STEM, OP_NODE, OP_DW_MULTI, OP_POOL, OP_LANEFUSE, OP_SE_POOL, OP_SE_SCALE, OP_SE_FUSED, OP_GCONV_SE, OP_DET, OP_LANE, OP_WAIT };

struct PlanOp {
    OpKind kind;
    int branch = 0;   // 0 = trunk (caller's stream); k > 0 = independent branch k, forked after the trunk
    ConvLaunch conv;  // OP_CONV (tensor maps pre-encoded)
    hn_stem_desc stem;
    hn_node_desc node;
    hn_dw_multi_desc dw_multi;
    hn_pool_desc pool;
    hn_lanefuse_desc lanefuse;
    hn_se_pool_desc se_pool;
    hn_se_scale_desc se_scale;
    hn_gconv_se_desc gconv_se;
    hn_det_desc det;
    hn_lane_desc lane;
    int wait_branch = 0;  // OP_WAIT: the branch whose completion this branch waits for here
};

static constexpr int kMaxBranches = 4;
static int g_branch_priority = 1;
extern "C" void hn_plan_set_branch_priority(int on) { g_branch_priority = on ? 1 : 0; }
struct hn_plan {
    std::vector<PlanOp*> ops;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int cur_branch = 0;  // branch given to the ops added next
    cudaStream_t side[kMaxBranches] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxBranches] = {nullptr, nullptr, nullptr, nullptr};
    ~hn_plan() {
        for (auto* o : ops) delete o;
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        for (int b = 0; b < kMaxBranches; ++b) {
            if (side[b]) cudaStreamDestroy(side[b]);
            if (ev_join[b]) cudaEventDestroy(ev_join[b]);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
    }
};

// Ops added after hn_plan_set_branch(p, k), k > 0, form branch k: they depend on the trunk (branch 0) ops added
// before them but not on other branches.  hn_plan_run forks them onto the plan's own streams after the trunk and joins
// them back into the caller's stream (inside a capture this becomes a forked graph): the three heads of the network
// read the same pyramid and write disjoint outputs, and at small batch each of their kernels fills only a few SMs.
extern "C" int hn_plan_set_branch(hn_plan* p, int branch) {
    HN_REQUIRE(p != nullptr && branch >= 0 && branch <= kMaxBranches, "bad branch %d", branch);
    HN_REQUIRE(branch >= p->cur_branch || branch == 0, "branches must be added in increasing order");
    p->cur_branch = branch;
    return HN_OK;
}

extern "C" int hn_plan_create(hn_plan** out) {
    HN_REQUIRE(out != nullptr, "null out");
    *out = new hn_plan();
    return HN_OK;
}
extern "C" int hn_plan_destroy(hn_plan* p) {
    delete p;
    return HN_OK;
}
extern "C" int hn_plan_size(const hn_plan* p) { return p ? (int)p->ops.size() : -1; }

#define PLAN_ADD(NAME, KIND, FIELD, DESC_T)                                  \
    extern "C" int hn_plan_add_##NAME(hn_plan* p, const DESC_T* d) {         \
        HN_REQUIRE(p != nullptr && d != nullptr, "null plan/desc");          \
        PlanOp* o = new PlanOp();                                            \
        o->kind = KIND;                                                      \
        o->branch = p->cur_branch;                                           \
        o->FIELD = *d;                                                       \
        p->ops.push_back(o);                                                 \
        return HN_OK;                                                        \
    }
PLAN_ADD(stem, OP_STEM, stem, hn_stem_desc)
PLAN_ADD(node, OP_NODE, node, hn_node_desc)
PLAN_ADD(dw_multi, OP_DW_MULTI, dw_multi, hn_dw_multi_desc)
PLAN_ADD(pool, OP_POOL, pool, hn_pool_desc)
PLAN_ADD(lanefuse, OP_LANEFUSE, lanefuse, hn_lanefuse_desc)
PLAN_ADD(se_pool, OP_SE_POOL, se_pool, hn_se_pool_desc)
PLAN_ADD(se_scale, OP_SE_SCALE, se_scale, hn_se_scale_desc)
PLAN_ADD(se_fused, OP_SE_FUSED, se_pool, hn_se_pool_desc)
PLAN_ADD(gconv_se, OP_GCONV_SE, gconv_se, hn_gconv_se_desc)
PLAN_ADD(det, OP_DET, det, hn_det_desc)
PLAN_ADD(lane, OP_LANE, lane, hn_lane_desc)

// A dependency between two branches: the ops added after it (in the current branch) start only once every op of branch
// `branch` -- which must have a smaller index, i.e. be enqueued earlier -- has finished.  The two detection towers run as
// two concurrent branches and the decode + NMS, which needs both, waits for the other tower.
extern "C" int hn_plan_add_wait(hn_plan* p, int branch) {
    HN_REQUIRE(p != nullptr, "null plan");
    HN_REQUIRE(branch >= 1 && branch < p->cur_branch, "plan: wait for branch %d from branch %d", branch, p->cur_branch);
    PlanOp* o = new PlanOp();
    o->kind = OP_WAIT;
    o->branch = p->cur_branch;
    o->wait_branch = branch;
    p->ops.push_back(o);
    return HN_OK;
}

extern "C" int hn_plan_add_conv(hn_plan* p, const hn_conv_desc* d) {
    HN_REQUIRE(p != nullptr && d != nullptr, "null plan/desc");
    PlanOp* o = new PlanOp();
    o->kind = OP_CONV;
    o->branch = p->cur_branch;
    int rc = hn_conv_prepare(d, &o->conv);
    if (rc) {
        delete o;
        return rc;
    }
    p->ops.push_back(o);
    return HN_OK;
}

static int op_launches(const PlanOp* o) {
    switch (o->kind) {
        case OP_DET: return hn_det_num_launches(&o->det);
        case OP_LANE: return 1;
        case OP_SE_POOL: return hn_se_pool_num_launches(&o->se_pool);
        case OP_WAIT: return 0;
        default: return 1;
    }
}

extern "C" int hn_plan_num_launches(const hn_plan* p) {
    if (!p) return -1;
    int n = 0;
    for (auto* o : p->ops) n += op_launches(o);
    return n;
}

static thread_local bool t_in_plan_run = false;
extern "C" int hn_plan_run_range(hn_plan* p, int first, int last, void* stream) {
    HN_REQUIRE(p != nullptr, "null plan");
    HN_REQUIRE(first >= 0 && last <= (int)p->ops.size() && first <= last, "bad op range [%d,%d)", first, last);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    for (int i = first; i < last; ++i) {
        PlanOp* o = p->ops[i];
        int rc = HN_OK;
        switch (o->kind) {
            case OP_CONV: rc = hn_conv_launch(&o->conv, s); break;
            case OP_STEM: rc = hn_stem_fwd(&o->stem, stream); break;
            case OP_NODE: rc = hn_node_fwd(&o->node, stream); break;
            case OP_DW_MULTI: rc = hn_dw_multi_fwd(&o->dw_multi, stream); break;
            case OP_POOL: rc = hn_pool_fwd(&o->pool, stream); break;
            case OP_LANEFUSE: rc = hn_lanefuse_fwd(&o->lanefuse, stream); break;
            case OP_SE_POOL: rc = hn_se_pool_fwd(&o->se_pool, stream); break;
            case OP_SE_SCALE: rc = hn_se_scale_fwd(&o->se_scale, stream); break;
            case OP_SE_FUSED: rc = hn_se_fused_fwd(&o->se_pool, stream); break;
            case OP_GCONV_SE: rc = hn_gconv_se_fwd(&o->gconv_se, stream); break;
            case OP_DET: rc = hn_det_decode_nms(&o->det, stream); break;
            case OP_LANE: rc = hn_lane_decode_nms(&o->lane, stream); break;
            case OP_WAIT:
                // only inside hn_plan_run, where the awaited branch has just been enqueued and its join event recorded in
                // the same (possibly capturing) context; a stand-alone run_range is sequential on one stream anyway
                if (t_in_plan_run && p->ev_join[o->wait_branch - 1]) HN_CHECK_CUDA(cudaStreamWaitEvent(s, p->ev_join[o->wait_branch - 1], 0));
                break;
        }
        if (rc) return rc;
    }
    return HN_OK;
}

static int plan_run_impl(hn_plan* p, void* stream);
extern "C" int hn_plan_run(hn_plan* p, void* stream) {
    t_in_plan_run = true;
    const int rc = plan_run_impl(p, stream);
    t_in_plan_run = false;
    return rc;
}
static int plan_run_impl(hn_plan* p, void* stream) {
    HN_REQUIRE(p != nullptr, "null plan");
    const int n = (int)p->ops.size();
    int first_branch_op = n;
    for (int i = 0; i < n; ++i)
        if (p->ops[i]->branch > 0) { first_branch_op = i; break; }
    if (first_branch_op == n) return hn_plan_run_range(p, 0, n, stream);
    for (int i = first_branch_op + 1; i < n; ++i)
        HN_REQUIRE(p->ops[i]->branch >= p->ops[i - 1]->branch, "plan: op %d (branch %d) after an op of branch %d", i,
                   p->ops[i]->branch, p->ops[i - 1]->branch);
    cudaStream_t main_s = reinterpret_cast<cudaStream_t>(stream);
    int rc = hn_plan_run_range(p, 0, first_branch_op, stream);
    if (rc) return rc;
    if (!p->ev_fork) HN_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    HN_CHECK_CUDA(cudaEventRecord(p->ev_fork, main_s));
    // branch 1 stays on the caller's stream, the others go to side streams
    int i = first_branch_op;
    bool used[kMaxBranches] = {false, false, false, false};
    while (i < n) {
        const int b = p->ops[i]->branch;
        int j = i;
        while (j < n && p->ops[j]->branch == b) ++j;
        if (b == p->ops[first_branch_op]->branch) {
            rc = hn_plan_run_range(p, i, j, stream);
            if (rc == HN_OK) {  // later branches may wait for this one (hn_plan_add_wait)
                if (!p->ev_join[b - 1]) HN_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_join[b - 1], cudaEventDisableTiming));
                HN_CHECK_CUDA(cudaEventRecord(p->ev_join[b - 1], main_s));
            }
        } else {
            const int k = b - 1;
            if (!p->side[k]) {
                // the side branches (detection towers + NMS, lanes) are the longer ones: give their CTAs precedence over
                // the caller-stream branch whenever an SM frees up
                int lo = 0, hi = 0;
                HN_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                HN_CHECK_CUDA(cudaStreamCreateWithPriority(&p->side[k], cudaStreamNonBlocking, g_branch_priority ? hi : lo));
            }
            if (!p->ev_join[k]) HN_CHECK_CUDA(cudaEventCreateWithFlags(&p->ev_join[k], cudaEventDisableTiming));
            HN_CHECK_CUDA(cudaStreamWaitEvent(p->side[k], p->ev_fork, 0));
            rc = hn_plan_run_range(p, i, j, p->side[k]);
            if (rc == HN_OK) HN_CHECK_CUDA(cudaEventRecord(p->ev_join[k], p->side[k]));
            used[k] = true;
        }
        if (rc) return rc;
        i = j;
    }
    for (int k = 0; k < kMaxBranches; ++k)
        if (used[k]) HN_CHECK_CUDA(cudaStreamWaitEvent(main_s, p->ev_join[k], 0));
    return HN_OK;
}

extern "C" int hn_plan_graph_capture(hn_plan* p, void* stream) {
    HN_REQUIRE(p != nullptr, "null plan");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (p->exec) {
        cudaGraphExecDestroy(p->exec);
        p->exec = nullptr;
    }
    if (p->graph) {
        cudaGraphDestroy(p->graph);
        p->graph = nullptr;
    }
    HN_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = hn_plan_run(p, stream);
    cudaError_t e = cudaStreamEndCapture(s, &p->graph);
    if (rc) return rc;
    HN_CHECK_CUDA(e);
    HN_CHECK_CUDA(cudaGraphInstantiate(&p->exec, p->graph, 0));
    return HN_OK;
}

extern "C" int hn_plan_graph_launch(hn_plan* p, void* stream) {
    HN_REQUIRE(p != nullptr && p->exec != nullptr, "plan has no captured graph");
    HN_CHECK_CUDA(cudaGraphLaunch(p->exec, reinterpret_cast<cudaStream_t>(stream)));
    return HN_OK;
}
