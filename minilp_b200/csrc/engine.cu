// minilp_b200 engine: device-resident revised-simplex state and the sm_100a kernels of the pivot path.
// Reference items replaced (file:line under /root/reference/src) are cited at each kernel / entry point.
// Compiled with -fmad=false: the reference (Rust) never contracts a*b+c, and every element-wise update
// here reproduces the reference's operation order exactly; only reductions (dot products, sums of
// squares, arg-min/max scans) are evaluated in a different — tree / chunked — order.
#include "minilp_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
extern "C" const char* mlp_last_error(void) { return g_err.c_str(); }
extern "C" void mlp_set_last_error(const char* msg) { g_err = msg ? msg : ""; }  // for the other translation units of the library
extern "C" const char* mlp_version(void) { return "minilp_b200 0.2 (sm_100a)"; }
extern "C" int mlp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

#define CU(x)                                                                                   \
  do {                                                                                          \
    cudaError_t err__ = (x);                                                                    \
    if (err__ != cudaSuccess) {                                                                 \
      set_err(std::string(#x) + ": " + cudaGetErrorString(err__) + " @" + std::to_string(__LINE__)); \
      return MLP_CUDA_ERROR;                                                                    \
    }                                                                                           \
  } while (0)
#define ST(x)                          \
  do {                                 \
    mlp_status st__ = (x);             \
    if (st__ != MLP_OK) return st__;   \
  } while (0)

static constexpr double EPS = 1e-8;  // solver.rs:12
#define FULLMASK 0xffffffffu

struct DevRes {  // small result block, mirrored in pinned host memory
  double f[8];
  long long i[6];
  int flags[4];  // [0] nonfinite, [1] singular
};
// Selection candidate exchanged between column shards: 64-byte header, followed in the exchange buffer by the
// candidate's column of [A|I] (m doubles).  Ordering: larger key wins, ties go to the smaller `tie`.
// tie = (position or variable index) << 32 | tie counts of the dual ratio test (pack_ties; 0 for pricing candidates): the
// high word is unique per candidate, so the low word never decides the order.
struct Cand {
  double key;
  long long tie;
  long long var;  // GLOBAL variable index, -1: none
  double f[5];    // primal: d, x_N, -, -, err   dual: coeff, d, x_N, pos, err
};
static_assert(sizeof(Cand) == 64, "Cand must be 64 bytes");

#include "kernels_common.cuh"
#include "chain_fused.cuh"
#include "sparse_build.cuh"
#include "dense_block.cuh"
#include "refresh_inverse.cuh"

// ------------------------------------------------------------------------------------------------ communicators
// The pivot path has ONE real exchange step per pivot: the arg-reduce of the per-shard pricing
// candidates, fused here with the distribution of the winner's column (one all-gather of 64 + 8m bytes per rank).
struct Comm {
  int rank = 0, world = 1;
  virtual ~Comm() {}
  virtual mlp_status allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st) = 0;
  virtual mlp_status broadcast(void* buf, size_t bytes, int root, cudaStream_t st) = 0;
};

// NCCL, loaded lazily so that single-GPU use has no dependency on libnccl.
namespace ncclapi {
typedef ncclResult_t (*GetUniqueId_t)(ncclUniqueId*);
typedef ncclResult_t (*CommInitRank_t)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*AllGather_t)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*Broadcast_t)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*CommDestroy_t)(ncclComm_t);
typedef const char* (*GetErrorString_t)(ncclResult_t);
static void* handle = nullptr;
static GetUniqueId_t GetUniqueId;
static CommInitRank_t CommInitRank;
static AllGather_t AllGather;
static Broadcast_t Broadcast;
static CommDestroy_t CommDestroy;
static GetErrorString_t GetErrorString;
static bool load() {
  if (handle) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_err(std::string("dlopen libnccl.so.2: ") + dlerror()); return false; }
  GetUniqueId = (GetUniqueId_t)dlsym(h, "ncclGetUniqueId");
  CommInitRank = (CommInitRank_t)dlsym(h, "ncclCommInitRank");
  AllGather = (AllGather_t)dlsym(h, "ncclAllGather");
  Broadcast = (Broadcast_t)dlsym(h, "ncclBroadcast");
  CommDestroy = (CommDestroy_t)dlsym(h, "ncclCommDestroy");
  GetErrorString = (GetErrorString_t)dlsym(h, "ncclGetErrorString");
  if (!GetUniqueId || !CommInitRank || !AllGather || !Broadcast || !CommDestroy || !GetErrorString) {
    set_err("libnccl is missing a required symbol");
    return false;
  }
  handle = h;
  return true;
}
}  // namespace ncclapi
#define NC(x)                                                                       \
  do {                                                                              \
    ncclResult_t r__ = (x);                                                         \
    if (r__ != ncclSuccess) {                                                       \
      set_err(std::string(#x) + ": " + ncclapi::GetErrorString(r__));               \
      return MLP_CUDA_ERROR;                                                        \
    }                                                                               \
  } while (0)

struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  ~NcclComm() override { if (comm) ncclapi::CommDestroy(comm); }
  mlp_status init(const void* id128, int rank_, int world_) {
    if (!ncclapi::load()) return MLP_INVALID;
    rank = rank_;
    world = world_;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NC(ncclapi::CommInitRank(&comm, world, id, rank));
    return MLP_OK;
  }
  mlp_status allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
    NC(ncclapi::AllGather(send, recv, bytes, ncclChar, comm, st));
    return MLP_OK;
  }
  mlp_status broadcast(void* buf, size_t bytes, int root, cudaStream_t st) override {
    NC(ncclapi::Broadcast(buf, buf, bytes, ncclChar, root, comm, st));
    return MLP_OK;
  }
};

// In-process group of logical shards driven by one host thread each (tests on a single GPU, or several GPUs of one
// process): rendezvous on a host barrier, data moves with cudaMemcpyAsync between the shards' device buffers.
struct LocalGroup {
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long gen = 0;
  std::vector<const void*> ptrs;
  explicit LocalGroup(int w) : world(w), ptrs(w, nullptr) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const long g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct LocalComm : Comm {
  LocalGroup* g = nullptr;
  mlp_status allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
    CU(cudaStreamSynchronize(st));
    g->ptrs[rank] = send;
    g->barrier();
    for (int r = 0; r < world; ++r)
      CU(cudaMemcpyAsync((char*)recv + (size_t)r * bytes, g->ptrs[r], bytes, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    g->barrier();
    return MLP_OK;
  }
  mlp_status broadcast(void* buf, size_t bytes, int root, cudaStream_t st) override {
    CU(cudaStreamSynchronize(st));
    g->ptrs[rank] = buf;
    g->barrier();
    if (rank != root) CU(cudaMemcpyAsync(buf, g->ptrs[root], bytes, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    g->barrier();
    return MLP_OK;
  }
};

// ------------------------------------------------------------------------------------------------ engine
// Per-stream scratch.  The pivot runs two chains concurrently: lane 0 carries the critical path
// (FTRAN of the entering column, BTRAN of alpha_q, the dense N^T v price-out, the reduced-cost update), lane 1 the
// latency-bound rest (ratio test, BTRAN of e_r, tableau-row price-out, FTRAN of rho, row updates, eta push).
struct Lane {
  cudaStream_t st = nullptr;
  double *xk = nullptr, *xk2 = nullptr;  // kcap: core right-hand side / solution
  double *tK = nullptr, *tK2 = nullptr;  // Kcap: eta scalars
  double* wm = nullptr;                   // m
  double* gpart = nullptr;                // TALL_MAXG x mld: column-group partials of the tall-skinny products
  double *gt_part_k = nullptr, *gt_part_K = nullptr;
  int32_t* seg_cnt = nullptr;
  double* seg_ss = nullptr;
  double* red_f = nullptr;      // 4096
  long long* red_i = nullptr;   // 4096
  unsigned* red_counter = nullptr;
  double* partial = nullptr;    // price partial sums: PR_MAXC x lda
  DevRes* d_res = nullptr;
  DevRes* h_res = nullptr;      // pinned
};

struct mlp_engine {
  int device = 0;
  cudaStream_t stream = nullptr;  // == lane[0].st
  Lane lane[2];
  // Column sharding: this engine owns structural columns [c0, c0+n) of the ng global ones; the m slack
  // variables are replicated on every shard.  LOCAL variable index: structural g -> g-c0, slack ng+i -> n+i.
  // Everything the ABI shows is GLOBAL.  world == 1: c0 = 0, n = ng.
  int64_t m = 0, n = 0, nt = 0, lda = 0, ng = 0, c0 = 0;
  int64_t mld = 0;  // row capacity (>= m): allocation length of every row-indexed array and leading dimension of Bcols / E;
                    // rows are appended by mlp_engine_add_row (Solver::add_constraint) without re-laying anything out
  Comm* comm = nullptr;
  int rank = 0, world = 1;
  int sm_count = 148;
  bool initialized = false;
  int enable_pse = 0, enable_dse = 0;

  double* A = nullptr;                                    // m x lda row-major, local column block (dense storage)
  // sparse storage (BASELINE config 4): CSR + CSC copies of A with 32-bit indices, as solver.rs:21-22 keeps both
  bool sparse = false;
  int64_t nnz = 0;
  int64_t nnz_loc = 0, sg0 = 0, sg1 = 0;                  // entries / segment range of this shard's column block [c0, c0+n)
  int64_t *csr_ptr = nullptr, *csc_ptr = nullptr;         // m+1, ng+1 (every shard holds the whole matrix)
  int32_t *csr_idx = nullptr, *csc_idx = nullptr;         // nnz
  double *csr_val = nullptr, *csc_val = nullptr;          // nnz
  std::vector<int64_t> h_csc_ptr;                         // host copy: column counts for LUFactors::nnz
  std::vector<int64_t> h_csr_ptr;                         // host copy of the CSR matrix: Solution::add_constraint appends a row and
  std::vector<int32_t> h_csr_idx;                         // rebuilds the CSC copy and the segment table from it, as the reference
  std::vector<double> h_csr_val;                          // rebuilds its CSR and CSC (solver.rs:598-610)
  int32_t *corevar = nullptr, *corepos = nullptr, *rowcore = nullptr;  // kcap, n, m: see k_ftran_finish_csr
  // compact row-major copy of the basic structural columns (rebuilt at every refactorization, k_ftran_finish_dcsr)
  int64_t* dcsr_ptr = nullptr;     // mld + 1
  int32_t* dcsr_idx = nullptr;     // dcsr_cap: core column of the entry
  double* dcsr_val = nullptr;      // dcsr_cap
  int64_t dcsr_cap = 0;
  int32_t* dcsr_hist = nullptr;    // DCSR_CHUNKS x mld
  int64_t* dcsr_cnt = nullptr;     // mld
  int64_t corevar_k = 0;                                  // entries of corevar currently marked in corepos
  // segment table of the CSC copy (<= CSC_SEG entries of one column per segment) and the core's slice of it
  int64_t nseg = 0;
  struct SegDesc* seg_desc = nullptr;  // nseg: (first entry, length, column) of every segment, one 16-byte load
  int32_t *seg_long = nullptr, *seg_short = nullptr;  // this shard's segments with more than / at most PR_CSC_LONG entries
  int64_t nseg_long = 0, nseg_short = 0;
  int32_t* seg_col = nullptr;      // nseg
  int64_t* seg_off = nullptr;      // nseg
  int64_t* col_seg = nullptr;      // n+1
  double* seg_sum = nullptr;       // nseg
  std::vector<int64_t> h_col_seg;  // host copy
  int32_t *cseg_id = nullptr, *cseg_first = nullptr;  // core segments (capacity cseg_cap) / kcap+1
  double* csum[2] = {nullptr, nullptr};                // per lane
  int64_t cseg_cap = 0, ncseg = 0;
  double *lo = nullptr, *hi = nullptr, *cobj = nullptr;  // ng+m, GLOBAL index, replicated
  double *d = nullptr, *gam = nullptr, *xnb = nullptr;   // n+m, local index
  uint8_t* vflag = nullptr;                               // n+m
  int32_t* vpos = nullptr;                                // n+m
  int32_t* bvar = nullptr;                                // m, GLOBAL variable ids
  double *xB = nullptr, *loB = nullptr, *hiB = nullptr, *w = nullptr, *rhs = nullptr;  // m
  double *alpha = nullptr, *rho = nullptr, *tau = nullptr, *vvec = nullptr;             // m
  double *work_m = nullptr, *work_mb = nullptr;                                          // m: BTRAN inputs of lane 0 / lane 1
  double* colq = nullptr;  // m: column of the entering variable; stored in the column cache by the pivot
  int64_t colq_var = -1;
  double *rc = nullptr, *helper = nullptr;  // n+m
  int32_t *list_idx = nullptr, *vlist_idx = nullptr;  // m: support of rho / of v = B^-T alpha_q
  double *list_val = nullptr, *vlist_val = nullptr;    // m
  double* scal = nullptr;   // device scalars: [0] max_step [1] |rho|^2 [2] |alpha|^2 [3] |v|^2 [4..6] objective parts
  int32_t* icnt = nullptr;  // device ints: [0] nnz rho [1] nnz alpha [2] nnz v
  DevRes* d_res = nullptr;  // == lane[0].d_res
  DevRes* h_res = nullptr;  // pinned, == lane[0].h_res
  // candidate exchange
  char *xsend = nullptr, *xrecv = nullptr;
  size_t xbytes = 0;
  Cand* h_cands = nullptr;  // pinned, world entries
  double* xred = nullptr;   // world * m doubles (vector sums across shards) / world scalars

  // dense LU of the basis
  int64_t k = 0, kcap = 0;
  int32_t *Jpos = nullptr, *Jslot = nullptr, *Rp = nullptr;  // kcap
  int32_t* rowcover = nullptr;                                 // m
  int32_t *lu_aff = nullptr, *lu_perm = nullptr;              // panel row-permutation records (192), in-place permutation (kcap)
  int32_t* lu_rcnt = nullptr;                                  // kcap (sparse storage): entries per core row, the reference's orig_row2elt_count
  unsigned long long* d_nnzcnt = nullptr;                      // [0] stored entries of the core before, [1] off-diagonal entries of L\U after the factorization
  double *Bcols = nullptr, *LUc = nullptr, *Cinv = nullptr;   // column cache m x kcap (slot-indexed); LU factors and (LU)^-1, kcap x kcap
  // eta file
  int64_t K = 0, Kcap = 0;
  double *E = nullptr, *Ginv = nullptr, *gK = nullptr;  // eta columns m x Kcap; (I+G)^-1 Kcap x Kcap; coupling row of the newest eta
  int32_t *etaR = nullptr, *etaPrev = nullptr, *etaHead = nullptr;
  int32_t* etaLast = nullptr;  // mld: index of the newest eta whose leaving row is r, -1 if none (head of k_eta_scatter's chain)
  uint8_t *touched = nullptr, *touched_new = nullptr;  // mld each: stored positions of the newest eta / of the column being priced in (k_touch_mark)
  // fused FTRAN -> BTRAN chain (chain_fused.cuh)
  int fused = 1;               // MLP_FUSED=0: separate kernels
  int fused_max = FZ_MAX;      // largest k / K that takes the fused chain (MLP_FUSED_MAX lowers it: tests of the hand-over)
  double* fz_scratch = nullptr;
  int32_t* fz_cta_cnt = nullptr;
  double* fz_cta_ss = nullptr;
  unsigned* fz_bar = nullptr;
  int64_t lu_nnz = 0;
  // product-form refresh of the core inverse (refresh_inverse.cuh; sparse storage): between two TRUE factorizations the
  // refactorizations the host asks for fold the eta file into C^-1 instead
  int64_t lu_every = 1 << 30;   // pivots between true factorizations (MLP_TUNE_LU_EVERY / MLP_LU_EVERY); 0 or 1: every refactorization is a true
                                // one.  Default: no count limit — the accuracy probe of every refresh (rf_tol) decides
  int64_t pivots_since_lu = 0;  // basis changes since the last true factorization
  std::vector<int32_t> h_Jpos_f, h_R_f;                       // core columns' positions / core rows of the factorized basis
  std::vector<int32_t> h_pos_core, h_row_core, h_rowcover_f;  // m each: position -> core column, row -> core row, row -> position of its basic slack (-1: none)
  std::vector<int32_t> h_eta_pos;    // position of every basis change since (= leaving position of every eta pushed since)
  std::vector<int64_t> h_eta_leave;  // ... and the variable that left it
  std::vector<int32_t> h_rc_old_rows, h_rc_old_vals;  // slack rows touched by incremental_sets and their rowcover values BEFORE it
  std::vector<int32_t> h_R_sorted;   // core rows of the factorized basis, ascending (h_R_f is in the factors' row order)
  bool inv_valid = false;            // Cinv holds the inverse of the core described by h_Jpos_f / h_R_f (false after a re-allocation)
  bool chg_complete = false;         // h_eta_pos / h_eta_leave list EVERY basis change since the last refactorization: the next one
                                     // may derive its index sets from the previous ones in O(k + K) instead of O(m)
  // pinned staging for the index arrays a refactorization uploads (two buffers, reused alternately behind an event)
  int32_t* stg_h[2] = {nullptr, nullptr};
  size_t stg_cap[2] = {0, 0};
  cudaEvent_t stg_ev[2] = {nullptr, nullptr};
  int stg_cur = 0;
  size_t stg_off = 0;
  int32_t* rf_map = nullptr;         // 3 kcap + 2 RF_MAXK: rowsrc | colsrc | jposn | etasrc | wrow
  double *rf_W = nullptr, *rf_T = nullptr, *rf_Ep = nullptr;  // RF_MAXK x kcap each
  double rf_worst_true = 0.0;               // (trace mode) the same probe right after true factorizations
  double rf_tol = 1e-10, rf_worst = 0.0;    // a refresh whose accuracy probe (k_rf_probe) exceeds rf_tol is redone as a true factorization
  double fill_true = 1.0;                   // off-diagonal entries of L\U per entry of the core, at the last true factorization

  // lane synchronisation (see "host side")
  int merge_small = 1;  // MLP_MERGE_SMALL=0: short eta files take the separate kernels too (A/B of k_eta_apply / k_unit_eta_t / k_eta_push)
  int use_pool = 1;     // MLP_POOL=0: growing arenas re-allocated with cudaMalloc / cudaFree instead of the stream-ordered pool
  int pdl = 1;          // MLP_PDL=0: ordinary launches (no programmatic dependent launch)
  int overlap = 1;      // MLP_OVERLAP=0: both lanes on one stream
  int async_pivot = 1;  // MLP_ASYNC_PIVOT=0: mlp_pivot always waits for the device
  int price_tma = 1;    // bulk-copy price-out kernel (MLP_PRICE_TMA=0: LDG kernel)
  int price_tile = 512; // its tile width in columns (MLP_PRICE_TILE; default: choose_price_tiling)
  int price_split = 1;  // column slices per item of the last, partial round (MLP_PRICE_SPLIT)
  int lane1_ldg = 1;    // lane 1 prices out with the LDG kernel while lane 0 runs the bulk-copy one (MLP_LANE1_LDG=0: both bulk-copy)
  int64_t inv_blocked_min = 2048;  // cores at least this large get their inverse by blocked substitution (MLP_INV_BLOCKED_MIN)
  int csc_stream = 1;      // CSC price-out loads the matrix evict-first (MLP_CSC_STREAM=0: default cache policy)
  int csc_grid = 148 * 5;  // CSC price-out: one full wave of resident CTAs (occupancy query at creation)
  int price_ctas = 6;   // resident price-out CTAs per SM (MLP_PRICE_CTAS); 6 = register-limited occupancy, measured 6.8 TB/s
                        // (4: 6.7, 3: 6.0, 2: 4.7 TB/s)
  cudaEvent_t s0_mark = nullptr, s1_mark = nullptr, ev_vbtran = nullptr, ev_win = nullptr;
  int64_t spec_var = -1;  // variable whose v = B^-T alpha_q / N^T v were computed ahead by mlp_ftran_col
  bool sel_valid = false; // xsend holds the pricing candidate of the CURRENT state (left by k_update_select)
  int64_t ftran_var = -1; // variable whose FTRAN (alpha, |alpha|^2) was queued right behind its selection
  Cand* d_win = nullptr;  // winner header of the last candidate exchange
  // peer-memory exchange (k_exchange_p2p): own buffer = 2 columns (by exchange parity) + 2 x world mailbox slots
  bool p2p = false;
  char* pbuf = nullptr;
  char* peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t pcol_bytes = 0, pbox_off = 0;
  unsigned long long xseq = 0;
  size_t smem_optin = 48 << 10;

  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // price-out timing, double-buffered by pivot parity: a pivot's events are read only after the NEXT selection has synced
  cudaEvent_t pev[2][2][2] = {};  // [slot][parity][begin/end]; slot 0 rho, 1 v
  bool ppending[2][2] = {};
  int32_t* h_mail = nullptr;      // pinned: [parity*4 + 0] nnz(rho), [parity*4 + 1] nnz(v) of the timed launches
  int64_t pivot_seq = 0;          // completed basis changes; parity = pivot_seq & 1
  int64_t alpha_nnz_host = -1;    // nnz(alpha_q) as read back with the ratio test, -1: not known on the host
  int64_t dual_row_host = -1;     // row whose basic value mlp_select_row_dual just returned (-1: none), and that value:
  double dual_row_val = 0.0;      // spares mlp_ratio_dual a device->host round trip per dual pivot
  int prof_on = 0;
  mlp_profile prof{};

  std::vector<int64_t> h_bvar;           // GLOBAL ids
  std::vector<int32_t> h_slot_of_row;    // cache slot of the structural basic variable at row r, or -1
  std::vector<int32_t> h_free_slots;
  std::vector<int32_t> h_pending_free;   // slots of columns that left the basis since the last refactor: the factors of
                                         // the refactor-time basis still read them (U = [D1; U2]), so they are recycled only then
  std::vector<int32_t> h_last_eta_of_row;
  mlp_counters cnt{};
  // MLP_REFACTOR_TRACE=1: wall time of every stage of refactor_impl, with a device sync after each (a diagnosis mode: the
  // syncs serialise what normally overlaps); printed to stderr when the engine is destroyed
  int refac_trace = 0;
  bool refac_in_pivot = false;
  double wait_ms = 0.0;  // (trace mode) host time blocked in the per-pivot device waits (fetch_res, winner header)
  int64_t waits = 0;
  std::vector<std::pair<const char*, double>> refac_stage;
  std::chrono::steady_clock::time_point refac_t;
  double refac_k_sum = 0.0;
};

// Kernel launch with the programmatic-dependent-launch attribute (see pdl_wait in kernels_common.cuh).
template <class... KArgs, class... Args>
static inline void launch_kernel(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  if (!pdl) {
    kern<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define LAUNCHS(e, st, kern, grid, block, smem, ...)                                        \
  do {                                                                                      \
    launch_kernel((e)->pdl != 0, kern, dim3(grid), dim3(block), (size_t)(smem), (st), __VA_ARGS__); \
    (e)->cnt.kernel_launches += 1;                                                          \
  } while (0)
#define LAUNCH(e, kern, grid, block, smem, ...) LAUNCHS(e, (e)->stream, kern, grid, block, smem, __VA_ARGS__)

// Arenas that grow DURING a solve (LU / eta / compact-row / core-segment arenas) are re-allocated from the device's
// stream-ordered pool (cudaMallocAsync / cudaFreeAsync, release threshold = keep everything): cudaFree synchronises the
// device and returns the pages to the driver — measured 7 to 130 ms per growth event on config 4, varying from run to run
// (profiles/r02e_refactor_trace.md) — while a pool free is just a stream operation and the next growth reuses the memory.
// PoolScope switches dev_alloc / dev_free of the calling thread to the pool for its lifetime.
static thread_local cudaStream_t g_pool_stream = nullptr;
struct PoolScope {
  cudaStream_t prev;
  explicit PoolScope(cudaStream_t st) : prev(g_pool_stream) { g_pool_stream = st; }
  ~PoolScope() { g_pool_stream = prev; }
};
template <class T> static mlp_status dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t err = g_pool_stream ? cudaMallocAsync((void**)p, count * sizeof(T), g_pool_stream) : cudaMalloc((void**)p, count * sizeof(T));
  if (err != cudaSuccess) {
    set_err(std::string("cudaMalloc: ") + cudaGetErrorString(err));
    return err == cudaErrorMemoryAllocation ? MLP_NOMEM : MLP_CUDA_ERROR;
  }
  return MLP_OK;
}
template <class T> static void dev_free(T*& p) {
  if (p) {
    if (g_pool_stream) cudaFreeAsync(p, g_pool_stream);
    else cudaFree(p);
  }
  p = nullptr;
}
static mlp_status h2d(mlp_engine* e, void* dst, const void* src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream));
  e->cnt.h2d_bytes += (int64_t)bytes;
  return MLP_OK;
}
static mlp_status d2h(mlp_engine* e, void* dst, const void* src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->cnt.d2h_bytes += (int64_t)bytes;
  return MLP_OK;
}
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// local index of a GLOBAL variable, -1 if the structural column lives on another shard
static inline int64_t to_local(const mlp_engine* e, int64_t g) {
  if (g >= e->ng) return e->n + (g - e->ng);
  return (g >= e->c0 && g < e->c0 + e->n) ? g - e->c0 : -1;
}
static inline int owner_of(const mlp_engine* e, int64_t g) {
  if (g >= e->ng) return e->rank;  // slack: replicated
  for (int r = 0; r < e->world; ++r) {
    int64_t b, en;
    mlp_shard_range(e->ng, e->world, r, &b, &en);
    if (g >= b && g < en) return r;
  }
  return -1;
}

#include "price_kernels.cuh"
#include "engine_kernels.cuh"

// ================================================================================================ host side
// Two lanes (streams).  API calls keep sequential semantics through two marks: s0_mark is recorded on lane 0 at the end of
// the non-speculative part of every lane-0 call, s1_mark on lane 1 at the end of every lane-1 call; a call on one lane
// first waits for the other lane's mark.  The only work that runs AHEAD of the marks is the speculative tail of
// mlp_ftran_col (v = B^-T alpha_q and the dense N^T v price-out), which touches nothing the lane-1 calls write except
// the eta file (guarded by ev_vbtran).
static mlp_status mark0(mlp_engine* e) { CU(cudaEventRecord(e->s0_mark, e->lane[0].st)); return MLP_OK; }
static mlp_status mark1(mlp_engine* e) { if (e->overlap) CU(cudaEventRecord(e->s1_mark, e->lane[1].st)); return MLP_OK; }
static mlp_status begin0(mlp_engine* e) { if (e->overlap) CU(cudaStreamWaitEvent(e->lane[0].st, e->s1_mark, 0)); return MLP_OK; }
static mlp_status begin1(mlp_engine* e) { if (e->overlap) CU(cudaStreamWaitEvent(e->lane[1].st, e->s0_mark, 0)); return MLP_OK; }
static mlp_status fetch_res(mlp_engine* e, Lane& ln) {
  CU(cudaMemcpyAsync(ln.h_res, ln.d_res, sizeof(DevRes), cudaMemcpyDeviceToHost, ln.st));
  const auto t0 = std::chrono::steady_clock::now();
  CU(cudaStreamSynchronize(ln.st));
  if (e->refac_trace) { e->wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); e->waits += 1; }
  e->cnt.d2h_bytes += (int64_t)sizeof(DevRes);
  return MLP_OK;
}
static int price_grid(const mlp_engine* e) { return e->sm_count * e->price_ctas; }
static void launch_price_tma(mlp_engine* e, cudaStream_t st, const int32_t* rows, const double* wts, const int32_t* count_ptr,
                             int fixed_count, double* partial) {
  const int tc = e->price_tile, sp = e->price_split;
  if (tc > 2048)
    LAUNCHS(e, st, k_price_partial_tma<8>, e->sm_count, TP_THREADS, TP_SMEM, e->A, e->lda, rows, wts, count_ptr, fixed_count, partial, tc, sp);
  else if (tc > 1024)
    LAUNCHS(e, st, k_price_partial_tma<4>, e->sm_count, TP_THREADS, TP_SMEM, e->A, e->lda, rows, wts, count_ptr, fixed_count, partial, tc, sp);
  else if (tc > 512)
    LAUNCHS(e, st, k_price_partial_tma<2>, e->sm_count, TP_THREADS, TP_SMEM, e->A, e->lda, rows, wts, count_ptr, fixed_count, partial, tc, sp);
  else
    LAUNCHS(e, st, k_price_partial_tma<1>, e->sm_count, TP_THREADS, TP_SMEM, e->A, e->lda, rows, wts, count_ptr, fixed_count, partial, tc, sp);
}

// out (local variable index) = N^T w over the listed rows (+ slack part), basic entries zeroed
static mlp_status price_list(mlp_engine* e, Lane& ln, const int32_t* rows, const double* wts, const int32_t* count_ptr,
                             int fixed_count, const double* slack_vals, double* out, int prof_slot = -1, bool finish = true) {
  const bool prof = e->prof_on && prof_slot >= 0;
  const int par = (int)(e->pivot_seq & 1);
  if (prof) CU(cudaEventRecord(e->pev[prof_slot][par][0], ln.st));
  if (e->sparse) {
    // slack_vals is the dense multiplier vector the list was compacted from
    double* ssum = &ln == &e->lane[0] ? e->seg_sum : e->seg_sum + e->nseg;
    if (e->csc_stream)
      LAUNCHS(e, ln.st, (k_price_csc_seg<0, true>), e->csc_grid, 256, 0, e->seg_desc, e->csc_idx, e->csc_val, e->seg_long, (int)e->nseg_long,
              e->seg_short, (int)e->nseg_short, slack_vals, ssum);
    else
      LAUNCHS(e, ln.st, (k_price_csc_seg<0, false>), e->csc_grid, 256, 0, e->seg_desc, e->csc_idx, e->csc_val, e->seg_long, (int)e->nseg_long,
              e->seg_short, (int)e->nseg_short, slack_vals, ssum);
    LAUNCHS(e, ln.st, k_price_csc_fin<0>, cdiv(e->nt, 256), 256, 0, e->col_seg, ssum, e->n, e->m, e->c0, slack_vals, e->vflag, out);
  } else {
    // Lane 1 (the tableau-row price-out, support <= k+1 rows) runs BESIDE lane 0's dense N^T v price-out.  The bulk-copy
    // kernel holds 194 KB of shared memory per SM, so a second instance cannot become resident until the first one has
    // finished (measured: the rho price-out waited 2.6 ms per pivot, and with it the whole tail of lane 1).  The LDG form
    // needs no shared-memory ring and slips in next to it; its partial sums are bit-identical.
    // (only while lane 0's N^T v of the same pivot is queued or running — spec_var — i.e. in the primal loop; in the dual
    // loop the tableau row is priced out before the entering column is known and has the GPU to itself)
    const bool beside = e->overlap && e->lane1_ldg && e->enable_pse && e->spec_var != -1 && &ln == &e->lane[1];
    if (e->price_tma && !beside)
      launch_price_tma(e, ln.st, rows, wts, count_ptr, fixed_count, ln.partial);
    else
      LAUNCHS(e, ln.st, k_price_partial<0>, beside ? e->sm_count * 2 : price_grid(e), PR_THREADS, 0, e->A, e->lda, rows, wts,
              count_ptr, fixed_count, ln.partial);
    // finish == false: the chunk partials are reduced by the consumer (k_update_select) instead
    if (finish)
      LAUNCHS(e, ln.st, k_price_finish, cdiv(e->nt, 256), 256, 0, ln.partial, count_ptr, fixed_count, e->lda, e->n, e->m,
              slack_vals, e->vflag, out, 0);
  }
  if (prof) {
    CU(cudaEventRecord(e->pev[prof_slot][par][1], ln.st));
    // support size of this launch, for the algorithmic byte count: lands in pinned memory in stream order
    if (count_ptr) CU(cudaMemcpyAsync(e->h_mail + par * 4 + prof_slot, count_ptr, sizeof(int32_t), cudaMemcpyDeviceToHost, ln.st));
    else e->h_mail[par * 4 + prof_slot] = fixed_count;
    e->ppending[prof_slot][par] = true;
  }
  return MLP_OK;
}
// fold the finished price-out timings of one parity into the profile (their events must have completed: called after a
// host sync that is ordered behind them)
static mlp_status collect_profile(mlp_engine* e, int par) {
  for (int slot = 0; slot < 2; ++slot) {
    if (!e->ppending[slot][par]) continue;
    e->ppending[slot][par] = false;
    float ms = 0.f;
    CU(cudaEventSynchronize(e->pev[slot][par][1]));
    CU(cudaEventElapsedTime(&ms, e->pev[slot][par][0], e->pev[slot][par][1]));
    const int64_t sz = e->h_mail[par * 4 + slot];
    const int64_t bytes = e->sparse ? 12 * e->nnz_loc + 8 * e->m + 8 * e->nt : 8 * e->n * sz + 8 * sz + 8 * e->n;
    if (slot == 0) { e->prof.price_rho_ms += ms; e->prof.price_rho_launches += 1; e->prof.price_rho_bytes += bytes; }
    else { e->prof.price_v_ms += ms; e->prof.price_v_launches += 1; e->prof.price_v_bytes += bytes; }
  }
  return MLP_OK;
}

// list / stats of a dense m-vector (see k_compact_count). idx == nullptr: stats only.
static void compact(mlp_engine* e, Lane& ln, const double* x, int32_t* idx, double* val, int32_t* count, double* sumsq,
                    const uint8_t* mask = nullptr) {
  const int m = (int)e->m, nseg = cdiv(m, CP_SEG);
  LAUNCHS(e, ln.st, k_compact_count, nseg, CP_SEG, 0, x, m, ln.seg_cnt, ln.seg_ss, ln.red_counter, count, sumsq, mask);
  if (idx) LAUNCHS(e, ln.st, k_compact_write, nseg, CP_SEG, 0, x, m, ln.seg_cnt, idx, val);
}
static int gemv_split(const mlp_engine* e, int rows, int cols) {
  // enough (column, row-slice) CTAs to fill every SM's thread slots (8 CTAs of 256 threads)
  int S = std::max(1, std::min(GT_MAXSPLIT, cdiv(8 * (int64_t)e->sm_count, cols)));
  return std::min(S, std::max(1, rows / 2048));
}
constexpr int TALL_MAXG = 16;
// column groups for the thread-per-row products: fill the SMs' thread slots, at least 32 columns per group
static int tall_groups(const mlp_engine* e, int rows, int cols) {
  const int want = cdiv((int64_t)e->sm_count * 2048, std::max(rows, 1));
  return std::max(1, std::min(std::min(TALL_MAXG, want), cols / 32));
}
static void gemv_t(mlp_engine* e, Lane& ln, const double* M, int64_t ld, int rows, int cols, const double* x, double* part,
                   const double* base, const int32_t* base_idx, double* out, int negate) {
  if (cols <= 0) return;
  const int S = gemv_split(e, rows, cols);
  LAUNCHS(e, ln.st, k_gemv_t_part, dim3((unsigned)cols, (unsigned)S), 256, 0, M, ld, rows, cols, x, part);
  LAUNCHS(e, ln.st, k_gemv_t_fin, cdiv(cols, 256), 256, 0, part, S, cols, base, base_idx, out, negate);
}

// BasisSolver::solve (solver.rs:1305-1319). rhs0: dense m-vector by constraint row (device). out: by basis position.
// mark: this is the FTRAN of an entering column — record the structural pattern of its result (k_touch_mark)
static mlp_status ftran(mlp_engine* e, Lane& ln, const double* rhs0, double* out, bool mark = false) {
  const int m = (int)e->m, k = (int)e->k, K = (int)e->K;
  // x = U^-1 L^-1 P a_R (lu.rs:92-93) as one product with the explicit inverse of the core
  if (k > 0) LAUNCHS(e, ln.st, k_mv_n<false>, cdiv(k, 32), 256, 0, e->Cinv, e->kcap, k, rhs0, e->Rp, ln.xk);
  const int Gk = tall_groups(e, m, k);
  const int GK = K > 0 ? tall_groups(e, m, K) : 1;
  // short eta file: its whole application is one launch (k_eta_apply) reading the LU part's result from lane scratch
  const bool eta_one = e->merge_small && K > 0 && K <= FE_MAXK && GK == 1 && rhs0 != ln.wm && out != ln.wm;
  double* mid = eta_one ? ln.wm : out;
  uint8_t* tn = mark ? e->touched_new : (uint8_t*)nullptr;  // structural pattern of an entering column's result (k_touch_mark)
  if (e->sparse) {
    LAUNCHS(e, ln.st, k_ftran_finish_dcsr, cdiv(std::max(m, k), 256), 256, 0, e->dcsr_ptr, e->dcsr_idx, e->dcsr_val, m, k, ln.xk, rhs0,
            e->rowcover, e->Jpos, mid, (const uint8_t*)e->touched, tn);
  } else if (Gk > 1) {
    LAUNCHS(e, ln.st, k_tall_part, dim3(cdiv(m, 256), Gk), 256, 0, e->Bcols, e->mld, m, k, ln.xk, e->Jslot, e->rowcover, ln.gpart, e->mld);
    LAUNCHS(e, ln.st, k_ftran_finish_parts, cdiv(std::max(m, k), 256), 256, 0, ln.gpart, Gk, e->mld, m, k, ln.xk, rhs0, e->rowcover,
            e->Jpos, mid, (const uint8_t*)e->touched, tn);
  } else
    LAUNCHS(e, ln.st, k_ftran_finish, cdiv(std::max(m, k), 256), 256, 0, e->Bcols, e->mld, m, k, ln.xk, rhs0, e->rowcover, e->Jpos,
            e->Jslot, mid, (const uint8_t*)e->touched, tn);
  if (eta_one) {
    LAUNCHS(e, ln.st, k_eta_apply, cdiv(m, 256), 256, 0, e->E, e->mld, m, K, e->Ginv, e->Kcap, e->etaR, (const double*)mid, out);
  } else if (K > 0) {  // eta file, solver.rs:1310-1316 in closed form: t = (I+G)^-1 alpha0[r], alpha -= E t
    LAUNCHS(e, ln.st, k_mv_n<true>, cdiv(K, 32), 256, 0, e->Ginv, e->Kcap, K, out, e->etaR, ln.tK);
    if (GK > 1) {
      LAUNCHS(e, ln.st, k_tall_part, dim3(cdiv(m, 256), GK), 256, 0, e->E, e->mld, m, K, ln.tK, (const int32_t*)nullptr,
              (const int32_t*)nullptr, ln.gpart, e->mld);
      LAUNCHS(e, ln.st, k_sub_parts, cdiv(m, 256), 256, 0, ln.gpart, GK, e->mld, m, out);
    } else
      LAUNCHS(e, ln.st, k_gemv_n_sub, cdiv(m, 256), 256, 0, e->E, e->mld, m, K, ln.tK, out);
  }
  return MLP_OK;
}

// BasisSolver::solve_transp (solver.rs:1322-1338). c: dense m-vector by basis position (device, DESTROYED).
// unit_row >= 0 tells that c == e_unit_row (the eta dot products degenerate to a row gather). out: by constraint row.
// gathered: tK already holds row unit_row of E (k_unit_and_gather)
static mlp_status btran(mlp_engine* e, Lane& ln, double* c, int unit_row, double* out, bool gathered = false, bool s_ready = false) {
  const int m = (int)e->m, k = (int)e->k, K = (int)e->K;
  if (K > 0) {  // etas in reverse, 1325-1333: u = E^T c, s = (I+G)^-T u, c[r_j] -= s_j
    if (s_ready) {}  // k_unit_eta_t left s in tK2
    else if (unit_row >= 0 && !gathered) LAUNCHS(e, ln.st, k_gather_row, cdiv(K, 256), 256, 0, e->E, e->mld, unit_row, K, ln.tK);
    else if (unit_row >= 0) {}
    else gemv_t(e, ln, e->E, e->mld, m, K, c, ln.gt_part_K, nullptr, nullptr, ln.tK, 0);
    if (!s_ready) LAUNCHS(e, ln.st, k_mv_t<true>, cdiv(K, 8), 256, 0, e->Ginv, e->Kcap, K, ln.tK, (const int32_t*)nullptr, ln.tK2);
    LAUNCHS(e, ln.st, k_eta_scatter, cdiv(K, 256), 256, 0, ln.tK2, e->etaR, e->etaPrev, e->etaHead, K, c);
  }
  LAUNCHS(e, ln.st, k_btran_start, cdiv(m, 256), 256, 0, c, e->rowcover, m, out, ln.wm);
  if (k > 0) {
    if (e->sparse) {
      double* cs = e->csum[&ln == &e->lane[0] ? 0 : 1];
      LAUNCHS(e, ln.st, k_core_rhs_seg, cdiv(e->ncseg, 8), 256, 0, e->csc_ptr, e->csc_idx, e->csc_val, e->seg_col, e->seg_off, e->cseg_id,
              (int)e->ncseg, ln.wm, cs);
      LAUNCHS(e, ln.st, k_core_rhs_fin, cdiv(k, 256), 256, 0, cs, e->cseg_first, k, c, e->Jpos, ln.xk);
    } else {
      const int S = gemv_split(e, m, k);
      LAUNCHS(e, ln.st, k_core_rhs_part, dim3((unsigned)k, (unsigned)S), 256, 0, e->Bcols, e->mld, m, k, e->Jslot, ln.wm, ln.gt_part_k);
      LAUNCHS(e, ln.st, k_gemv_t_fin, cdiv(k, 256), 256, 0, ln.gt_part_k, S, k, c, e->Jpos, ln.xk, 1);
    }
    // y = L^-T U^-T rhs (lu_factors_transp, lu.rs:108-115) = (C^-1)^T rhs, scattered to the core's constraint rows
    LAUNCHS(e, ln.st, k_mv_t<false>, cdiv(k, 8), 256, 0, e->Cinv, e->kcap, k, ln.xk, e->Rp, out);
  }
  return MLP_OK;
}

#include "refactor_host.cuh"

// The exchange step: every shard's candidate header + candidate column are all-gathered, the winner is chosen with the
// reference's tie rule on the host, and its column becomes colq.  world == 1: no collective, same code path.
static mlp_status ftran(mlp_engine* e, Lane& ln, const double* rhs0, double* out, bool mark);
static mlp_status se_helper(mlp_engine* e, int64_t var);
static bool chain_fusable(const mlp_engine* e);
static mlp_status chain_fused(mlp_engine* e, int64_t var);
static void compact(mlp_engine* e, Lane& ln, const double* x, int32_t* idx, double* val, int32_t* count, double* sumsq,
                    const uint8_t* mask);
static constexpr int64_t VAR_PENDING = -2;

// want_alpha_nnz (dual loop): the host also waits for nnz(alpha_q) of the winner's FTRAN, so that mlp_pivot can return
// without a device round trip of its own (the primal loop gets that count with the ratio test's read-back).
static mlp_status exchange_candidates(mlp_engine* e, Cand* winner, bool want_alpha_nnz = false) {
  const int m = (int)e->m;
  Lane& l0 = e->lane[0];
  // single shard: the candidate's column goes straight into colq and its header is the winner
  double* dst = e->world > 1 ? (double*)(e->xsend + sizeof(Cand)) : e->colq;
  if (e->world > 1 && e->p2p) dst = (double*)(e->pbuf + (size_t)((e->xseq + 1) & 1) * e->pcol_bytes);  // this exchange's parity
  Cand* win = e->world > 1 ? nullptr : e->d_win;
  if (e->sparse) {
    CU(cudaMemsetAsync(dst, 0, (size_t)m * sizeof(double), e->stream));
    LAUNCH(e, k_load_col_csc, 4, 256, 0, e->csc_ptr, e->csc_idx, e->csc_val, e->ng, (int64_t)-1, (const Cand*)e->xsend, dst, win);
  } else {
    LAUNCH(e, k_cand_load_col, cdiv(m, 256), 256, 0, e->A, e->lda, e->n, e->c0, e->ng, m, (const Cand*)e->xsend, dst, win);
  }
  if (e->world > 1 && e->p2p) {
    PeerTable pt;
    for (int r = 0; r < 8; ++r) pt.base[r] = e->peer_base[r];
    e->xseq += 1;
    LAUNCH(e, k_exchange_p2p, 64, 256, 0, pt, e->rank, e->world, e->xseq, (int)(e->xseq & 1), (const Cand*)e->xsend, m,
           e->pcol_bytes, e->pbox_off, e->colq, e->d_win);
  } else if (e->world > 1) {
    ST(e->comm->allgather(e->xsend, e->xrecv, e->xbytes, e->stream));
    LAUNCH(e, k_pick_winner, cdiv(m, 256), 256, 0, e->xrecv, e->xbytes, e->world, m, e->colq, e->d_win);
  }
  CU(cudaMemcpyAsync(e->h_cands, e->d_win, sizeof(Cand), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev_win, e->stream));
  // Both callers continue with calc_col_coeffs of the winner (solver.rs:750, 532): queue that FTRAN — and, with primal
  // steepest edge, the v / N^T v chain — now, before the host has even seen which variable won.
  e->spec_var = -1;
  if (chain_fusable(e)) ST(chain_fused(e, VAR_PENDING));
  else {
    ST(ftran(e, l0, e->colq, e->alpha, true));
    compact(e, l0, e->alpha, nullptr, nullptr, e->icnt + 1, e->scal + 2, e->touched_new);
    ST(mark0(e));
    if (want_alpha_nnz && e->async_pivot) {  // header + nnz(alpha_q) in one wait, still ahead of the steepest-edge tail
      CU(cudaMemcpyAsync(e->h_mail + 6, e->icnt + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaEventRecord(e->ev_win, e->stream));
    } else want_alpha_nnz = false;
    if (e->enable_pse && e->overlap) ST(se_helper(e, VAR_PENDING));
  }
  if (chain_fusable(e)) want_alpha_nnz = false;
  e->alpha_nnz_host = -1;
  {
    const auto t0 = std::chrono::steady_clock::now();
    CU(cudaEventSynchronize(e->ev_win));  // the header only (dual loop: and the FTRAN), not the chain queued behind it
    if (e->refac_trace) { e->wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); e->waits += 1; }
  }
  if (e->prof_on) ST(collect_profile(e, (int)((e->pivot_seq & 1) ^ 1)));  // the previous pivot is complete by now
  e->cnt.d2h_bytes += (int64_t)sizeof(Cand);
  *winner = e->h_cands[0];
  if (want_alpha_nnz && winner->var >= 0) { e->alpha_nnz_host = e->h_mail[6]; e->cnt.d2h_bytes += 4; }
  if (winner->f[4] == 2.0) { set_err("device-side rendezvous timed out (peer-memory exchange: a rank stopped participating; or a grid barrier of the fused chain)"); return MLP_CUDA_ERROR; }
  if (winner->f[4] != 0.0) { set_err("non-finite steepest-edge norm"); return MLP_NONFINITE; }
  e->colq_var = e->ftran_var = winner->var;
  if (e->spec_var == VAR_PENDING) e->spec_var = winner->var;
  return MLP_OK;
}

// make colq hold the column of GLOBAL variable var on every shard
static mlp_status fetch_column(mlp_engine* e, int64_t var) {
  if (e->colq_var == var) return MLP_OK;
  const int m = (int)e->m;
  const int64_t lv = to_local(e, var);
  if (e->sparse) {  // every shard holds the whole matrix: no broadcast
    CU(cudaMemsetAsync(e->colq, 0, (size_t)m * sizeof(double), e->stream));
    LAUNCH(e, k_load_col_csc, 4, 256, 0, e->csc_ptr, e->csc_idx, e->csc_val, e->ng, var, (const Cand*)nullptr, e->colq, (Cand*)nullptr);
  } else if (lv >= 0) LAUNCH(e, k_load_col, cdiv(m, 256), 256, 0, e->A, e->lda, e->n, m, lv, e->colq);
  if (!e->sparse && e->world > 1 && var < e->ng) ST(e->comm->broadcast(e->colq, (size_t)m * sizeof(double), owner_of(e, var), e->stream));
  e->colq_var = var;
  return MLP_OK;
}

// update_primal_sq_norms' bulk part on lane 0: v = B^-T alpha_q (solver.rs:1114), helper = N^T v (1117-1132).
// Depends only on alpha_q, so mlp_ftran_col starts it ahead of the ratio test; the eta file may change only after ev_vbtran.
static mlp_status se_helper(mlp_engine* e, int64_t var) {
  Lane& l0 = e->lane[0];
  CU(cudaMemcpyAsync(e->work_m, e->alpha, (size_t)e->m * 8, cudaMemcpyDeviceToDevice, l0.st));
  ST(btran(e, l0, e->work_m, -1, e->vvec));
  CU(cudaEventRecord(e->ev_vbtran, l0.st));
  compact(e, l0, e->vvec, e->vlist_idx, e->vlist_val, e->icnt + 2, e->scal + 3);
  ST(price_list(e, l0, e->vlist_idx, e->vlist_val, e->icnt + 2, 0, e->vvec, e->helper, 1, false));
  e->spec_var = var;
  return MLP_OK;
}

// The same chain as ftran + compact + se_helper, as one cooperative launch (chain_fused.cuh) followed by the price-out.
static bool chain_fusable(const mlp_engine* e) {
  return e->fused && !e->sparse && e->enable_pse && e->overlap && e->k <= e->fused_max && e->K <= e->fused_max;
}
static int fz_slices(const mlp_engine* e, int cols) {
  // about two (column, slice) units per warp of the grid; a slice keeps >= 512 rows.  Depends only on m, cols and the SM
  // count, so every shard of a column-sharded engine reduces in the same order.
  const int64_t W = (int64_t)e->sm_count * FZ_WARPS;
  const int want = cdiv(2 * W, std::max(cols, 1));
  const int cap = (int)std::max<int64_t>(1, std::min<int64_t>(FZ_MAXS, e->m / 512));
  return std::max(1, std::min(want, cap));
}
static mlp_status chain_fused(mlp_engine* e, int64_t var) {
  Lane& l0 = e->lane[0];
  ChainArgs a;
  a.m = (int)e->m; a.k = (int)e->k; a.K = (int)e->K;
  a.S_k = fz_slices(e, a.k); a.S_K = fz_slices(e, a.K);
  a.mld = e->mld; a.kcap = e->kcap; a.Kcap = e->Kcap;
  a.Cinv = e->Cinv; a.Ginv = e->Ginv; a.Bcols = e->Bcols; a.E = e->E; a.colq = e->colq;
  a.Rp = e->Rp; a.Jpos = e->Jpos; a.Jslot = e->Jslot; a.rowcover = e->rowcover;
  a.etaR = e->etaR; a.etaPrev = e->etaPrev; a.etaLast = e->etaLast;
  a.alpha = e->alpha; a.vvec = e->vvec; a.cov = l0.wm;
  double* z = e->fz_scratch;
  a.px = z; z += FZ_G * FZ_MAX;
  a.pt = z; z += FZ_G * FZ_MAX;
  a.pu = z; z += FZ_MAXS * FZ_MAX;
  a.pr = z; z += FZ_MAXS * FZ_MAX;
  a.uK = z; z += FZ_MAX;
  a.sK = z; z += FZ_MAX;
  a.rk = z;
  a.cta_cnt = e->fz_cta_cnt; a.cta_ss = e->fz_cta_ss;
  a.seg_cnt = l0.seg_cnt; a.seg_ss = l0.seg_ss;
  a.vidx = e->vlist_idx; a.vval = e->vlist_val;
  a.icnt = e->icnt; a.scal = e->scal;
  a.bar = e->fz_bar; a.flags = e->d_res->flags;
  a.touched = e->touched; a.touched_new = e->touched_new;
  void* args[] = {&a};
  CU(cudaLaunchCooperativeKernel((const void*)k_chain_primal, dim3((unsigned)e->sm_count), dim3(FZ_T), args, 0, l0.st));
  e->cnt.kernel_launches += 1;
  ST(mark0(e));                              // alpha_q, nnz(alpha_q), |alpha_q|^2 are final: lane 1 may start the ratio test
  CU(cudaEventRecord(e->ev_vbtran, l0.st));  // ... and the eta file has been read for the last time
  ST(price_list(e, l0, e->vlist_idx, e->vlist_val, e->icnt + 2, 0, e->vvec, e->helper, 1, false));
  e->spec_var = var;
  return MLP_OK;
}

// Map every rank's exchange buffer into this process (CUDA IPC over NVLink).  All ranks agree on the outcome through an
// all-gather of their success flags; on any failure everyone keeps the NCCL all-gather path.
static void setup_p2p(mlp_engine* e, NcclComm* nc) {
  if (const char* v = getenv("MLP_P2P")) if (atoi(v) == 0) return;
  if (e->world > 8) return;
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
  static_assert(sizeof(Msg) == 128, "Msg layout");
  const size_t col = ((size_t)e->mld * sizeof(double) + 255) / 256 * 256;
  const size_t total = 2 * col + 2 * (size_t)e->world * P2P_SLOT;
  Msg mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.ok = 1;
  if (cudaMalloc((void**)&e->pbuf, total) != cudaSuccess) { cudaGetLastError(); e->pbuf = nullptr; mine.ok = 0; }
  if (mine.ok && cudaMemset(e->pbuf, 0, total) != cudaSuccess) mine.ok = 0;
  if (mine.ok && cudaIpcGetMemHandle(&mine.h, e->pbuf) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
  Msg* d_msgs = nullptr;
  std::vector<Msg> all((size_t)e->world);
  bool coll_ok = cudaMalloc((void**)&d_msgs, sizeof(Msg) * (e->world + 1)) == cudaSuccess;
  if (coll_ok) {
    cudaMemcpyAsync(d_msgs + e->world, &mine, sizeof(Msg), cudaMemcpyHostToDevice, e->stream);
    coll_ok = nc->allgather(d_msgs + e->world, d_msgs, sizeof(Msg), e->stream) == MLP_OK;
    cudaMemcpyAsync(all.data(), d_msgs, sizeof(Msg) * e->world, cudaMemcpyDeviceToHost, e->stream);
    coll_ok = cudaStreamSynchronize(e->stream) == cudaSuccess && coll_ok;
    cudaFree(d_msgs);
  }
  bool ok = coll_ok;
  for (int r = 0; ok && r < e->world; ++r) ok = all[(size_t)r].ok != 0;
  int opened = 1;
  if (ok) {
    for (int r = 0; r < e->world; ++r) {
      if (r == e->rank) { e->peer_base[r] = e->pbuf; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
      e->peer_base[r] = (char*)p;
    }
  }
  // second agreement round: did everyone manage to map everyone?
  int* d_flag = nullptr;
  std::vector<int> flags((size_t)e->world, 0);
  int my_flag = ok && opened;
  if (coll_ok && cudaMalloc((void**)&d_flag, sizeof(int) * (e->world + 1)) == cudaSuccess) {
    cudaMemcpyAsync(d_flag + e->world, &my_flag, sizeof(int), cudaMemcpyHostToDevice, e->stream);
    const bool g = nc->allgather(d_flag + e->world, d_flag, sizeof(int), e->stream) == MLP_OK;
    cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * e->world, cudaMemcpyDeviceToHost, e->stream);
    const bool s2 = cudaStreamSynchronize(e->stream) == cudaSuccess;
    cudaFree(d_flag);
    bool every = g && s2;
    for (int r = 0; every && r < e->world; ++r) every = flags[(size_t)r] != 0;
    e->p2p = every;
  }
  e->pcol_bytes = col;
  e->pbox_off = 2 * col;
  if (!e->p2p) {
    for (int r = 0; r < 8; ++r) {
      if (e->peer_base[r] && e->peer_base[r] != e->pbuf) cudaIpcCloseMemHandle(e->peer_base[r]);
      e->peer_base[r] = nullptr;
    }
  }
}

static void destroy_engine(mlp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  for (int l = 0; l < 2; ++l) if (e->lane[l].st) cudaStreamSynchronize(e->lane[l].st);
  refac_report(e);
  dev_free(e->csr_ptr); dev_free(e->csc_ptr); dev_free(e->csr_idx); dev_free(e->csc_idx); dev_free(e->csr_val); dev_free(e->csc_val);
  dev_free(e->corevar); dev_free(e->corepos); dev_free(e->rowcore);
  dev_free(e->seg_col); dev_free(e->seg_off); dev_free(e->col_seg); dev_free(e->seg_sum); dev_free(e->cseg_id); dev_free(e->cseg_first);
  dev_free(e->seg_desc); dev_free(e->seg_long); dev_free(e->seg_short);
  dev_free(e->rf_map); dev_free(e->rf_W); dev_free(e->rf_T); dev_free(e->rf_Ep);
  for (int c = 0; c < 2; ++c) {
    if (e->stg_h[c]) cudaFreeHost(e->stg_h[c]);
    if (e->stg_ev[c]) cudaEventDestroy(e->stg_ev[c]);
  }
  dev_free(e->dcsr_ptr); dev_free(e->dcsr_idx); dev_free(e->dcsr_val); dev_free(e->dcsr_hist); dev_free(e->dcsr_cnt);
  dev_free(e->csum[0]); dev_free(e->csum[1]);
  dev_free(e->A); dev_free(e->lo); dev_free(e->hi); dev_free(e->cobj); dev_free(e->d); dev_free(e->gam); dev_free(e->xnb);
  dev_free(e->vflag); dev_free(e->vpos); dev_free(e->bvar); dev_free(e->xB); dev_free(e->loB); dev_free(e->hiB); dev_free(e->w);
  dev_free(e->rhs); dev_free(e->alpha); dev_free(e->rho); dev_free(e->tau); dev_free(e->vvec); dev_free(e->work_m);
  dev_free(e->work_mb); dev_free(e->colq); dev_free(e->rc); dev_free(e->helper); dev_free(e->list_idx); dev_free(e->list_val);
  dev_free(e->vlist_idx); dev_free(e->vlist_val); dev_free(e->scal); dev_free(e->icnt);
  dev_free(e->xsend); dev_free(e->xrecv); dev_free(e->xred); dev_free(e->d_win);
  for (int r = 0; r < 8; ++r) if (e->peer_base[r] && e->peer_base[r] != e->pbuf) cudaIpcCloseMemHandle(e->peer_base[r]);
  dev_free(e->pbuf);
  dev_free(e->rowcover); dev_free(e->Jpos); dev_free(e->Jslot); dev_free(e->Rp); dev_free(e->Bcols); dev_free(e->LUc); dev_free(e->Cinv);
  dev_free(e->lu_aff); dev_free(e->lu_perm); dev_free(e->lu_rcnt); dev_free(e->d_nnzcnt);
  dev_free(e->E); dev_free(e->Ginv); dev_free(e->gK); dev_free(e->etaR); dev_free(e->etaPrev); dev_free(e->etaHead);
  dev_free(e->etaLast); dev_free(e->touched); dev_free(e->touched_new); dev_free(e->fz_scratch); dev_free(e->fz_cta_cnt); dev_free(e->fz_cta_ss); dev_free(e->fz_bar);
  for (int l = 0; l < 2; ++l) {
    Lane& ln = e->lane[l];
    dev_free(ln.xk); dev_free(ln.xk2); dev_free(ln.tK); dev_free(ln.tK2); dev_free(ln.wm); dev_free(ln.gpart); dev_free(ln.gt_part_k);
    dev_free(ln.gt_part_K); dev_free(ln.seg_cnt); dev_free(ln.seg_ss); dev_free(ln.red_f); dev_free(ln.red_i);
    dev_free(ln.red_counter); dev_free(ln.partial); dev_free(ln.d_res);
    if (ln.h_res) cudaFreeHost(ln.h_res);
  }
  if (e->h_cands) cudaFreeHost(e->h_cands);
  for (int i = 0; i < 4; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int q = 0; q < 2; ++q) if (e->pev[i][j][q]) cudaEventDestroy(e->pev[i][j][q]);
  if (e->h_mail) cudaFreeHost(e->h_mail);
  if (e->s0_mark) cudaEventDestroy(e->s0_mark);
  if (e->s1_mark) cudaEventDestroy(e->s1_mark);
  if (e->ev_vbtran) cudaEventDestroy(e->ev_vbtran);
  if (e->ev_win) cudaEventDestroy(e->ev_win);
  delete e->comm;
  if (e->lane[1].st && e->lane[1].st != e->lane[0].st) cudaStreamDestroy(e->lane[1].st);
  if (e->lane[0].st) cudaStreamDestroy(e->lane[0].st);
  delete e;
}

static mlp_status create_engine(int device, int64_t m, int64_t ng, int rank, int world, Comm* comm, mlp_engine** out,
                                bool sparse = false, int64_t mld_override = 0);
static mlp_status create_engine(int device, int64_t m, int64_t ng, int rank, int world, Comm* comm, mlp_engine** out,
                                bool sparse, int64_t mld_override) {
  *out = nullptr;
  if (m <= 0 || ng <= 0 || m > 0x7fffffff || ng + m > 0x7fffffff || world < 1 || rank < 0 || rank >= world) {
    set_err("bad dimensions");
    delete comm;
    return MLP_INVALID;
  }
  if (mlp_device_count() <= device) {
    set_err("no CUDA device: the engine has no CPU fallback");
    delete comm;
    return MLP_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); delete comm; return MLP_CUDA_ERROR; }
  mlp_engine* e = new mlp_engine();
  struct Guard {  // every early return below releases the engine, its streams and the communicator
    mlp_engine* e;
    ~Guard() { if (e) destroy_engine(e); }
  } guard{e};
  e->device = device;
  e->sparse = sparse;
  e->comm = comm;
  e->rank = rank;
  e->world = world;
  int64_t c0, c1;
  mlp_shard_range(ng, world, rank, &c0, &c1);
  e->m = m; e->ng = ng; e->c0 = c0; e->n = c1 - c0; e->nt = e->n + m;
  e->lda = std::max<int64_t>(16, (e->n + 15) / 16 * 16);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  e->sm_count = prop.multiProcessorCount;
  if (const char* v = getenv("MLP_OVERLAP")) e->overlap = atoi(v) != 0;
  if (const char* v = getenv("MLP_PDL")) e->pdl = atoi(v) != 0;
  if (const char* v = getenv("MLP_POOL")) e->use_pool = atoi(v) != 0;
  if (const char* v = getenv("MLP_MERGE_SMALL")) e->merge_small = atoi(v) != 0;
  if (const char* v = getenv("MLP_REFRESH_TOL")) e->rf_tol = atof(v);
  if (e->use_pool) {  // keep what the growing arenas free (see PoolScope)
    cudaMemPool_t mp = nullptr;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&mp, device) != cudaSuccess || cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) {
      cudaGetLastError();
      e->use_pool = 0;
    }
  }
  if (const char* v = getenv("MLP_REFACTOR_TRACE")) e->refac_trace = atoi(v);
  if (const char* v = getenv("MLP_LU_EVERY")) e->lu_every = std::max<int64_t>(0, atoll(v));
  if (const char* v = getenv("MLP_CSC_STREAM")) e->csc_stream = atoi(v) != 0;
  if (const char* v = getenv("MLP_INV_BLOCKED_MIN")) e->inv_blocked_min = std::max<int64_t>(1, atoll(v));
  if (const char* v = getenv("MLP_ASYNC_PIVOT")) e->async_pivot = atoi(v) != 0;
  if (const char* v = getenv("MLP_PRICE_CTAS")) e->price_ctas = std::max(1, std::min(8, atoi(v)));
  if (const char* v = getenv("MLP_PRICE_TMA")) e->price_tma = atoi(v) != 0;
  if (const char* v = getenv("MLP_LANE1_LDG")) e->lane1_ldg = atoi(v) != 0;
  if (const char* v = getenv("MLP_FUSED")) e->fused = atoi(v) != 0;
  if (const char* v = getenv("MLP_FUSED_MAX")) e->fused_max = std::max(0, std::min(FZ_MAX, atoi(v)));
  {  // the fused chain needs a cooperative launch of one CTA per SM
    int nb = 0;
    if (!prop.cooperativeLaunch || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_chain_primal, FZ_T, 0) != cudaSuccess || nb < 1) {
      cudaGetLastError();
      e->fused = 0;
    }
  }
  {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_price_csc_seg<0, true>, 256, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 4; }
    e->csc_grid = e->sm_count * nb;
  }
  choose_price_tiling(e->lda, m, e->sm_count, &e->price_tile, &e->price_split);
  if (const char* v = getenv("MLP_PRICE_TILE")) {
    const int t = atoi(v);
    if (t >= 128 && t <= 4096 && t % 64 == 0) { e->price_tile = t; e->price_split = 1; }
  }
  if (const char* v = getenv("MLP_PRICE_SPLIT")) {
    const int t = atoi(v);
    if ((t == 1 || t == 2 || t == 4) && e->price_tile / t >= 128 && (e->price_tile / t) % 16 == 0) e->price_split = t;
  }
  CU(cudaFuncSetAttribute(k_price_partial_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM));
  CU(cudaFuncSetAttribute(k_price_partial_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM));
  CU(cudaFuncSetAttribute(k_price_partial_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM));
  CU(cudaFuncSetAttribute(k_price_partial_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM));
  {  // lane 1 carries short latency-bound kernels that must slip in beside the price-out: highest priority
    int lo_p = 0, hi_p = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    CU(cudaStreamCreateWithPriority(&e->lane[0].st, cudaStreamNonBlocking, lo_p));
    if (e->overlap) CU(cudaStreamCreateWithPriority(&e->lane[1].st, cudaStreamNonBlocking, hi_p));
    else e->lane[1].st = e->lane[0].st;
    e->stream = e->lane[0].st;
  }
  e->smem_optin = std::min<size_t>((size_t)prop.sharedMemPerBlockOptin, (size_t)200 << 10);
  CU(cudaFuncSetAttribute(k_core_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_optin));
  CU(cudaFuncSetAttribute(k_lu_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_optin));
  mlp_status st = MLP_OK;
  auto A = [&](mlp_status s) { if (st == MLP_OK) st = s; };
  // Row capacity: Solution::add_constraint / add_gomory_cut (lib.rs:368-423) append rows; every row-indexed array is
  // allocated for mld rows so that appending one costs O(n), not a re-layout.
  e->mld = m + std::max<int64_t>(64, m / 8);
  if (const char* v = getenv("MLP_ROW_RESERVE")) e->mld = m + std::max<int64_t>(0, atoll(v));
  if (mld_override >= m) e->mld = mld_override;
  const int64_t ml = e->mld, ntc = e->n + ml, gt = ng + ml;
  if (!sparse) A(dev_alloc(&e->A, (size_t)ml * e->lda));
  A(dev_alloc(&e->lo, gt)); A(dev_alloc(&e->hi, gt)); A(dev_alloc(&e->cobj, gt));
  A(dev_alloc(&e->d, ntc)); A(dev_alloc(&e->gam, ntc)); A(dev_alloc(&e->xnb, ntc));
  A(dev_alloc(&e->vflag, ntc)); A(dev_alloc(&e->vpos, ntc)); A(dev_alloc(&e->bvar, ml));
  A(dev_alloc(&e->xB, ml)); A(dev_alloc(&e->loB, ml)); A(dev_alloc(&e->hiB, ml)); A(dev_alloc(&e->w, ml)); A(dev_alloc(&e->rhs, ml));
  A(dev_alloc(&e->alpha, ml)); A(dev_alloc(&e->rho, ml)); A(dev_alloc(&e->tau, ml)); A(dev_alloc(&e->vvec, ml));
  A(dev_alloc(&e->work_m, ml)); A(dev_alloc(&e->work_mb, ml)); A(dev_alloc(&e->colq, ml));
  A(dev_alloc(&e->rc, ntc)); A(dev_alloc(&e->helper, ntc));
  A(dev_alloc(&e->list_idx, ml)); A(dev_alloc(&e->list_val, ml)); A(dev_alloc(&e->vlist_idx, ml)); A(dev_alloc(&e->vlist_val, ml));
  A(dev_alloc(&e->scal, 16)); A(dev_alloc(&e->icnt, 16)); A(dev_alloc(&e->d_nnzcnt, 2 + 2 * RF_PROBE));
  for (int l = 0; l < 2; ++l) {
    Lane& ln = e->lane[l];
    A(dev_alloc(&ln.wm, ml));
    A(dev_alloc(&ln.gpart, (size_t)16 * ml));
    A(dev_alloc(&ln.partial, (size_t)PR_MAXC * e->lda));
    const size_t nred = std::max<size_t>(4096, (size_t)cdiv(ntc, 256) + 1);  // k_update_select: one partial per 256 variables
    A(dev_alloc(&ln.red_f, nred)); A(dev_alloc(&ln.red_i, nred)); A(dev_alloc(&ln.red_counter, 4));
    A(dev_alloc(&ln.seg_cnt, (size_t)cdiv(ml, CP_SEG) + 1)); A(dev_alloc(&ln.seg_ss, (size_t)cdiv(ml, CP_SEG) + 1));
    A(dev_alloc(&ln.d_res, 1));
  }
  e->d_res = e->lane[0].d_res;
  A(dev_alloc(&e->rowcover, ml));
  A(dev_alloc(&e->etaLast, ml)); A(dev_alloc(&e->touched, ml)); A(dev_alloc(&e->touched_new, ml));
  A(dev_alloc(&e->fz_scratch, (size_t)(2 * FZ_G + 2 * FZ_MAXS + 3) * FZ_MAX));
  A(dev_alloc(&e->fz_cta_cnt, (size_t)e->sm_count)); A(dev_alloc(&e->fz_cta_ss, (size_t)e->sm_count)); A(dev_alloc(&e->fz_bar, 4));
  e->xbytes = sizeof(Cand) + (size_t)ml * sizeof(double);
  A(dev_alloc(&e->d_win, 1));
  A(dev_alloc(&e->xsend, e->xbytes)); A(dev_alloc(&e->xrecv, e->xbytes * world)); A(dev_alloc(&e->xred, (size_t)world * ml + 64));
  if (st != MLP_OK) return st;
  for (int l = 0; l < 2; ++l) CU(cudaHostAlloc((void**)&e->lane[l].h_res, sizeof(DevRes), cudaHostAllocDefault));
  e->h_res = e->lane[0].h_res;
  CU(cudaHostAlloc((void**)&e->h_cands, sizeof(Cand) * world, cudaHostAllocDefault));
  for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&e->ev[i]));
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int q = 0; q < 2; ++q) CU(cudaEventCreate(&e->pev[i][j][q]));
  CU(cudaHostAlloc((void**)&e->h_mail, 8 * sizeof(int32_t), cudaHostAllocDefault));
  std::memset(e->h_mail, 0, 8 * sizeof(int32_t));
  CU(cudaEventCreateWithFlags(&e->s0_mark, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->s1_mark, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->ev_vbtran, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->ev_win, cudaEventDisableTiming));
  if (!sparse) CU(cudaMemsetAsync(e->A, 0, (size_t)e->mld * e->lda * sizeof(double), e->stream));
  for (int l = 0; l < 2; ++l) {
    CU(cudaMemsetAsync(e->lane[l].red_counter, 0, 4 * sizeof(unsigned), e->stream));
    CU(cudaMemsetAsync(e->lane[l].d_res, 0, sizeof(DevRes), e->stream));
  }
  CU(cudaMemsetAsync(e->etaLast, 0xff, (size_t)ml * sizeof(int32_t), e->stream));
  CU(cudaMemsetAsync(e->touched, 0, (size_t)ml, e->stream));
  CU(cudaMemsetAsync(e->touched_new, 0, (size_t)ml, e->stream));
  CU(cudaMemsetAsync(e->fz_bar, 0, 4 * sizeof(unsigned), e->stream));
  CU(cudaMemsetAsync(e->gam, 0, ntc * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->helper, 0, ntc * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->rc, 0, ntc * sizeof(double), e->stream));
  CU(cudaMemsetAsync(e->xsend, 0, e->xbytes, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->h_bvar.assign(m, 0);
  e->h_slot_of_row.assign(m, -1);
  e->h_last_eta_of_row.assign(m, -1);
  guard.e = nullptr;
  *out = e;
  return MLP_OK;
}

extern "C" {

// ---------------------------------------------------------------------------------- sharding helpers (host only)
void mlp_shard_range(int64_t n, int32_t world, int32_t rank, int64_t* begin, int64_t* end) {
  // contiguous blocks of whole 16-column units (128-byte row segments), sizes differ by at most one unit
  const int64_t units = (n + 15) / 16;
  const int64_t b = units * rank / world, e = units * (rank + 1) / world;
  *begin = std::min<int64_t>(b * 16, n);
  *end = std::min<int64_t>(e * 16, n);
}
int32_t mlp_reduce_candidates(const double* scores, const int64_t* pos, const int64_t* vars, int32_t world) {
  int32_t best = -1;
  for (int32_t r = 0; r < world; ++r) {
    if (vars[r] < 0) continue;
    if (best < 0 || scores[r] > scores[best] || (scores[r] == scores[best] && pos[r] < pos[best])) best = r;
  }
  return best;
}
mlp_status mlp_local_group_create(int32_t world, void** out) {
  if (world < 1) return MLP_INVALID;
  *out = new LocalGroup(world);
  return MLP_OK;
}
void mlp_local_group_destroy(void* g) { delete (LocalGroup*)g; }
mlp_status mlp_nccl_get_unique_id(void* out128) {
  if (!ncclapi::load()) return MLP_INVALID;
  ncclUniqueId id;
  NC(ncclapi::GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return MLP_OK;
}

mlp_status mlp_engine_create_dense(int device, int64_t m, int64_t n, mlp_engine** out) {
  return create_engine(device, m, n, 0, 1, nullptr, out);
}
// (Re)build everything the engine derives from the host CSR copy (h_csr_*): the device CSR arrays, and ON THE DEVICE
// (sparse_build.cuh) the CSC copy — CsMat::to_csc (solver.rs:253, 610), rows ascending within a column as sparse.rs:230-269
// produces them — and the segment table of the CSC copy.  m = number of rows of the CSR copy.  MLP_HOST_TRANSPOSE=1 builds
// the CSC copy with the host counting transpose instead (kept for the equality test).
static mlp_status sparse_upload(mlp_engine* e, int64_t m) {
  const int64_t n = e->ng, nnz = (int64_t)e->h_csr_idx.size();  // the WHOLE matrix on every shard
  const int64_t* row_ptr = e->h_csr_ptr.data();
  const int32_t* col_idx = e->h_csr_idx.data();
  const double* vals = e->h_csr_val.data();
  e->nnz = nnz;
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  dev_free(e->csr_ptr); dev_free(e->csr_idx); dev_free(e->csr_val); dev_free(e->csc_ptr); dev_free(e->csc_idx); dev_free(e->csc_val);
  dev_free(e->seg_col); dev_free(e->seg_off); dev_free(e->col_seg); dev_free(e->seg_sum); dev_free(e->seg_desc);
  dev_free(e->seg_long); dev_free(e->seg_short);
  mlp_status st = MLP_OK;
  auto A = [&](mlp_status s2) { if (st == MLP_OK) st = s2; };
  A(dev_alloc(&e->csr_ptr, m + 1)); A(dev_alloc(&e->csr_idx, nnz)); A(dev_alloc(&e->csr_val, nnz));
  A(dev_alloc(&e->csc_ptr, n + 1)); A(dev_alloc(&e->csc_idx, nnz)); A(dev_alloc(&e->csc_val, nnz));
  A(dev_alloc(&e->col_seg, n + 1));
  if (st != MLP_OK) return st;
  ST(h2d(e, e->csr_ptr, row_ptr, (m + 1) * sizeof(int64_t)));
  ST(h2d(e, e->csr_idx, col_idx, nnz * sizeof(int32_t)));
  ST(h2d(e, e->csr_val, vals, nnz * sizeof(double)));
  e->h_csc_ptr.assign((size_t)n + 1, 0);
  e->h_col_seg.assign((size_t)n + 1, 0);
  bool host_transpose = false;
  if (const char* v = getenv("MLP_HOST_TRANSPOSE")) host_transpose = atoi(v) != 0;
  if (host_transpose) {
    std::vector<int64_t>& cptr = e->h_csc_ptr;
    for (int64_t t = 0; t < nnz; ++t) cptr[(size_t)col_idx[t] + 1] += 1;
    for (int64_t j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
    std::vector<int32_t> cidx((size_t)nnz);
    std::vector<double> cval((size_t)nnz);
    std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
    for (int64_t i = 0; i < m; ++i)
      for (int64_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) {
        const int64_t d = fill[col_idx[t]]++;
        cidx[d] = (int32_t)i;
        cval[d] = vals[t];
      }
    for (int64_t j = 0; j < n; ++j) {
      const int64_t c = cptr[j + 1] - cptr[j];
      e->h_col_seg[j + 1] = e->h_col_seg[j] + (c == 0 ? 1 : (c + CSC_SEG - 1) / CSC_SEG);
    }
    ST(h2d(e, e->csc_ptr, cptr.data(), (n + 1) * sizeof(int64_t)));
    ST(h2d(e, e->csc_idx, cidx.data(), nnz * sizeof(int32_t)));
    ST(h2d(e, e->csc_val, cval.data(), nnz * sizeof(double)));
    ST(h2d(e, e->col_seg, e->h_col_seg.data(), (n + 1) * sizeof(int64_t)));
    CU(cudaStreamSynchronize(e->stream));  // host vectors go out of scope
  } else {
    // chunks of consecutive rows: about two CTAs per SM for the fill, bounded by 256 MB of per-chunk column counters
    int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(2 * e->sm_count, m), ((int64_t)256 << 20) / (4 * std::max<int64_t>(n, 1))));
    const int rpc = (int)((m + chunks - 1) / chunks);
    chunks = (int)((m + rpc - 1) / rpc);
    int32_t* hist = nullptr;
    int64_t *cnt = nullptr, *segs = nullptr;
    A(dev_alloc(&hist, (size_t)chunks * n)); A(dev_alloc(&cnt, n)); A(dev_alloc(&segs, n));
    if (st == MLP_OK) {
      cudaMemsetAsync(hist, 0, (size_t)chunks * n * sizeof(int32_t), e->stream);
      LAUNCH(e, k_t_hist, cdiv(m * 32, 256), 256, 0, e->csr_ptr, e->csr_idx, m, n, rpc, hist, (const int32_t*)nullptr);
      LAUNCH(e, k_t_colscan, cdiv(n, 256), 256, 0, hist, n, chunks, cnt, segs, CSC_SEG);
      LAUNCH(e, k_scan_excl, 1, 1024, 0, cnt, n, e->csc_ptr);
      LAUNCH(e, k_scan_excl, 1, 1024, 0, segs, n, e->col_seg);
      LAUNCH(e, k_t_fill, chunks, 256, 0, e->csr_ptr, e->csr_idx, e->csr_val, m, n, rpc, hist, e->csc_ptr, e->csc_idx, e->csc_val,
             (const int32_t*)nullptr);
      // the host keeps the two pointer arrays: column counts for LUFactors::nnz, segment ranges for the core's segment list
      A(d2h(e, e->h_csc_ptr.data(), e->csc_ptr, (n + 1) * sizeof(int64_t)));
      A(d2h(e, e->h_col_seg.data(), e->col_seg, (n + 1) * sizeof(int64_t)));
    }
    dev_free(hist); dev_free(cnt); dev_free(segs);
    if (st != MLP_OK) return st;
    if (e->h_csc_ptr[(size_t)n] != nnz) { set_err("sparse upload: device transpose lost entries"); return MLP_CUDA_ERROR; }
  }
  e->nseg = e->h_col_seg[(size_t)n];
  e->sg0 = e->h_col_seg[(size_t)e->c0];
  e->sg1 = e->h_col_seg[(size_t)(e->c0 + e->n)];
  e->nnz_loc = e->h_csc_ptr[(size_t)(e->c0 + e->n)] - e->h_csc_ptr[(size_t)e->c0];
  A(dev_alloc(&e->seg_col, e->nseg)); A(dev_alloc(&e->seg_off, e->nseg)); A(dev_alloc(&e->seg_desc, e->nseg));
  A(dev_alloc(&e->seg_sum, 2 * e->nseg));  // one set per lane
  if (st != MLP_OK) return st;
  LAUNCH(e, k_t_segs, cdiv(n, 256), 256, 0, e->csc_ptr, e->col_seg, n, CSC_SEG, e->seg_col, e->seg_off, (int4*)e->seg_desc);
  {  // work lists of the price-out over this shard's column block (host: O(segments))
    std::vector<int32_t> lg, sh;
    for (int64_t j = e->c0; j < e->c0 + e->n; ++j) {
      const int64_t cnt = e->h_csc_ptr[(size_t)j + 1] - e->h_csc_ptr[(size_t)j];
      for (int64_t sg = e->h_col_seg[(size_t)j], t = 0; sg < e->h_col_seg[(size_t)j + 1]; ++sg, ++t) {
        const int64_t len = std::min<int64_t>(CSC_SEG, cnt - t * CSC_SEG);
        (len > PR_CSC_LONG ? lg : sh).push_back((int32_t)sg);
      }
    }
    e->nseg_long = (int64_t)lg.size();
    e->nseg_short = (int64_t)sh.size();
    A(dev_alloc(&e->seg_long, lg.size())); A(dev_alloc(&e->seg_short, sh.size()));
    if (st != MLP_OK) return st;
    if (!lg.empty()) ST(h2d(e, e->seg_long, lg.data(), lg.size() * sizeof(int32_t)));
    if (!sh.empty()) ST(h2d(e, e->seg_short, sh.data(), sh.size() * sizeof(int32_t)));
    CU(cudaStreamSynchronize(e->stream));  // host vectors go out of scope
  }
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) { set_err("sparse upload failed"); return MLP_CUDA_ERROR; }
  return MLP_OK;
}
// the CSC copy as the engine holds it (parity tests of the on-device transpose)
mlp_status mlp_engine_download_csc(mlp_engine* e, int64_t* col_ptr, int32_t* row_idx, double* vals) {
  if (!e || !e->sparse) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(begin0(e));
  if (col_ptr) ST(d2h(e, col_ptr, e->csc_ptr, (e->ng + 1) * sizeof(int64_t)));
  if (row_idx) ST(d2h(e, row_idx, e->csc_idx, e->nnz * sizeof(int32_t)));
  if (vals) ST(d2h(e, vals, e->csc_val, e->nnz * sizeof(double)));
  return MLP_OK;
}
// communicator of a column-sharded engine (nullptr for a single shard)
static mlp_status make_comm(int device, int32_t rank, int32_t world, int32_t comm_kind, const void* comm_arg, Comm** out) {
  *out = nullptr;
  if (world <= 1) return MLP_OK;
  if (mlp_device_count() <= device) { set_err("no CUDA device: the engine has no CPU fallback"); return MLP_NO_DEVICE; }
  CU(cudaSetDevice(device));
  if (comm_kind == MLP_COMM_NCCL) {
    NcclComm* c = new NcclComm();
    mlp_status st = c->init(comm_arg, rank, world);
    if (st != MLP_OK) { delete c; return st; }
    *out = c;
  } else if (comm_kind == MLP_COMM_LOCAL) {
    LocalComm* c = new LocalComm();
    c->g = (LocalGroup*)comm_arg;
    c->rank = rank;
    c->world = world;
    if (!c->g || c->g->world != world) { delete c; set_err("local group size mismatch"); return MLP_INVALID; }
    *out = c;
  } else { set_err("unknown comm kind"); return MLP_INVALID; }
  return MLP_OK;
}
mlp_status mlp_engine_create_sparse_sharded(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                                            const int32_t* col_idx, const double* vals, int32_t rank, int32_t world,
                                            int32_t comm_kind, const void* comm_arg, mlp_engine** out) {
  *out = nullptr;
  if (!row_ptr || !col_idx || !vals || nnz < 0 || m <= 0 || n <= 0 || row_ptr[0] != 0 || row_ptr[m] != nnz) {
    set_err("create_sparse: bad CSR");
    return MLP_INVALID;
  }
  for (int64_t i = 0; i < m; ++i)
    if (row_ptr[i + 1] < row_ptr[i]) { set_err("create_sparse: row_ptr not monotone"); return MLP_INVALID; }
  for (int64_t t = 0; t < nnz; ++t)
    if (col_idx[t] < 0 || col_idx[t] >= n) { set_err("create_sparse: column index out of range"); return MLP_INVALID; }
  for (int64_t i = 0; i < m; ++i)  // CsVec::new (lib.rs:279): sorted, no repeated index — the device transpose relies on it
    for (int64_t t = row_ptr[i] + 1; t < row_ptr[i + 1]; ++t)
      if (col_idx[t] <= col_idx[t - 1]) { set_err("create_sparse: columns must be strictly ascending within a row"); return MLP_INVALID; }
  Comm* comm = nullptr;
  ST(make_comm(device, rank, world, comm_kind, comm_arg, &comm));
  mlp_engine* e = nullptr;
  ST(create_engine(device, m, n, rank, world, comm, &e, true));
  e->h_csr_ptr.assign(row_ptr, row_ptr + m + 1);
  e->h_csr_idx.assign(col_idx, col_idx + nnz);
  e->h_csr_val.assign(vals, vals + nnz);
  mlp_status st = MLP_OK;
  auto A = [&](mlp_status s2) { if (st == MLP_OK) st = s2; };
  A(dev_alloc(&e->corepos, n)); A(dev_alloc(&e->rowcore, e->mld));
  if (st == MLP_OK && cudaMemsetAsync(e->corepos, 0xff, n * sizeof(int32_t), e->stream) != cudaSuccess) st = MLP_CUDA_ERROR;
  A(sparse_upload(e, m));
  if (st != MLP_OK) { destroy_engine(e); return st; }
  if (world > 1 && comm_kind == MLP_COMM_NCCL) setup_p2p(e, (NcclComm*)comm);  // best effort: falls back to the all-gather
  *out = e;
  return MLP_OK;
}
mlp_status mlp_engine_create_sparse(int device, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr, const int32_t* col_idx,
                                    const double* vals, mlp_engine** out) {
  return mlp_engine_create_sparse_sharded(device, m, n, nnz, row_ptr, col_idx, vals, 0, 1, MLP_COMM_NONE, nullptr, out);
}
mlp_status mlp_engine_create_dense_sharded(int device, int64_t m, int64_t n_global, int32_t rank, int32_t world,
                                           int32_t comm_kind, const void* comm_arg, mlp_engine** out) {
  *out = nullptr;
  Comm* comm = nullptr;
  ST(make_comm(device, rank, world, comm_kind, comm_arg, &comm));
  ST(create_engine(device, m, n_global, rank, world, comm, out));
  if (world > 1 && comm_kind == MLP_COMM_NCCL) setup_p2p(*out, (NcclComm*)comm);  // best effort: falls back to the all-gather
  return MLP_OK;
}
void mlp_engine_destroy(mlp_engine* e) { destroy_engine(e); }
int32_t mlp_engine_exchange_kind(mlp_engine* e) {
  if (!e || e->world <= 1) return 0;
  if (e->p2p) return 3;
  return dynamic_cast<NcclComm*>(e->comm) ? 1 : 2;
}
mlp_status mlp_engine_local_range(mlp_engine* e, int64_t* begin, int64_t* end) {
  if (!e) return MLP_INVALID;
  *begin = e->c0;
  *end = e->c0 + e->n;
  return MLP_OK;
}

// rows_host: nrows x src_cols row-major; src_cols == n_global (full rows: this shard's slice is taken) or == local width
static mlp_status upload_rows_impl(mlp_engine* e, int64_t row0, int64_t nrows, const double* rows_host, bool local) {
  if (!e || row0 < 0 || nrows < 0 || row0 + nrows > e->m) { set_err("upload_rows: range"); return MLP_INVALID; }
  if (e->sparse) { set_err("upload_rows: the engine holds a sparse matrix (given at creation)"); return MLP_INVALID; }
  CU(cudaSetDevice(e->device));
  if (e->n == 0 || nrows == 0) return MLP_OK;
  const double* src = local ? rows_host : rows_host + e->c0;
  const size_t spitch = (local ? e->n : e->ng) * sizeof(double);
  CU(cudaMemcpy2DAsync(e->A + row0 * e->lda, e->lda * sizeof(double), src, spitch, e->n * sizeof(double), (size_t)nrows,
                       cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->cnt.h2d_bytes += nrows * e->n * (int64_t)sizeof(double);
  return MLP_OK;
}
mlp_status mlp_engine_upload_rows(mlp_engine* e, int64_t row0, int64_t nrows, const double* rows_host) {
  return upload_rows_impl(e, row0, nrows, rows_host, false);
}
mlp_status mlp_engine_upload_local_rows(mlp_engine* e, int64_t row0, int64_t nrows, const double* rows_local) {
  return upload_rows_impl(e, row0, nrows, rows_local, true);
}

mlp_status mlp_engine_init_state(mlp_engine* e, const mlp_init_state* st) {
  if (!e || !st) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int64_t m = e->m, n = e->n, nt = e->nt, ng = e->ng, gt = ng + m;
  ST(h2d(e, e->lo, st->orig_var_mins, gt * 8));
  ST(h2d(e, e->hi, st->orig_var_maxs, gt * 8));
  ST(h2d(e, e->cobj, st->orig_obj_coeffs, gt * 8));
  ST(h2d(e, e->rhs, st->orig_rhs, m * 8));
  std::vector<double> d(nt, 0.0), xnb(nt, 0.0), gam(nt, 0.0);
  std::vector<uint8_t> fl(nt, 0);
  std::vector<int32_t> pos(nt, 0), bvar(m);
  e->h_bvar.assign(m, 0);
  for (int64_t c = 0; c < ng; ++c) {  // nb_vars has one entry per structural column count (global)
    const int64_t g = st->nb_vars[c];
    if (g < 0 || g >= gt) { set_err("init_state: nb_vars"); return MLP_INVALID; }
    const int64_t v = to_local(e, g);
    if (v < 0) continue;
    d[v] = st->nb_var_obj_coeffs[c];
    xnb[v] = st->nb_var_vals[c];
    fl[v] = st->nb_var_states[c] & (MLP_AT_MIN | MLP_AT_MAX | MLP_FIXED);
    pos[v] = (int32_t)c;
    if (st->primal_edge_sq_norms) gam[v] = st->primal_edge_sq_norms[c];
  }
  bool any_structural_basic = false;
  for (int64_t r = 0; r < m; ++r) {
    const int64_t g = st->basic_vars[r];
    if (g < 0 || g >= gt) { set_err("init_state: basic_vars"); return MLP_INVALID; }
    if (g < ng) any_structural_basic = true;
    const int64_t v = to_local(e, g);
    if (v >= 0) { fl[v] = MLP_BASIC; pos[v] = (int32_t)r; }
    bvar[r] = (int32_t)g;
    e->h_bvar[r] = g;
  }
  e->enable_pse = st->enable_primal_steepest_edge;
  e->enable_dse = st->enable_dual_steepest_edge;
  e->chg_complete = false;  // a whole new basis: the next refactorization builds its index sets from scratch
  e->h_Jpos_f.clear();
  ST(h2d(e, e->d, d.data(), nt * 8));
  ST(h2d(e, e->xnb, xnb.data(), nt * 8));
  ST(h2d(e, e->gam, gam.data(), nt * 8));
  ST(h2d(e, e->vflag, fl.data(), nt));
  ST(h2d(e, e->vpos, pos.data(), nt * 4));
  ST(h2d(e, e->bvar, bvar.data(), m * 4));
  ST(h2d(e, e->loB, st->basic_var_mins, m * 8));
  ST(h2d(e, e->hiB, st->basic_var_maxs, m * 8));
  if (st->basic_var_vals) ST(h2d(e, e->xB, st->basic_var_vals, m * 8));
  if (st->dual_edge_sq_norms) ST(h2d(e, e->w, st->dual_edge_sq_norms, m * 8));
  CU(cudaStreamSynchronize(e->stream));
  e->spec_var = -1;
  e->sel_valid = false;
  e->dual_row_host = -1;
  if (!st->basic_var_vals) {
    double* part = e->world > 1 ? e->work_m : e->xred;
    if (e->sparse) LAUNCH(e, k_row_dot_csr, cdiv(m * 32, 256), 256, 0, e->csr_ptr, e->csr_idx, e->csr_val, m, e->c0, n, e->xnb, part);
    else LAUNCH(e, k_row_dot, (unsigned)m, 256, 0, e->A, e->lda, n, e->xnb, part);
    if (e->world > 1) ST(e->comm->allgather(part, e->xred, (size_t)m * sizeof(double), e->stream));
    LAUNCH(e, k_init_basic_vals, cdiv(m, 256), 256, 0, e->xred, e->world, (int)m, e->rhs, e->xB);
  }
  if (!st->dual_edge_sq_norms) LAUNCH(e, k_fill, cdiv(m, 256), 256, 0, e->w, m, 1.0);
  if (e->enable_pse && !st->primal_edge_sq_norms) {
    // |a_j|^2 + 1 (solver.rs:297-299): all m rows, unit weights
    if (e->sparse)
    {
      LAUNCH(e, (k_price_csc_seg<1, true>), e->csc_grid, 256, 0, e->seg_desc, e->csc_idx, e->csc_val, e->seg_long, (int)e->nseg_long,
             e->seg_short, (int)e->nseg_short, (const double*)nullptr, e->seg_sum);
      LAUNCH(e, k_price_csc_fin<1>, cdiv(nt, 256), 256, 0, e->col_seg, e->seg_sum, n, m, e->c0, (const double*)nullptr, e->vflag, e->gam);
    }
    else {
    LAUNCH(e, k_price_partial<1>, price_grid(e), PR_THREADS, 0, e->A, e->lda, (const int32_t*)nullptr, (const double*)nullptr,
           (const int32_t*)nullptr, (int32_t)m, e->lane[0].partial);
    LAUNCH(e, k_price_finish, cdiv(nt, 256), 256, 0, e->lane[0].partial, (const int32_t*)nullptr, (int32_t)m, e->lda, n, m,
           (const double*)nullptr, e->vflag, e->gam, 1);
    }
  }
  // column cache: an initial basis with structural columns (warm start) fetches them one by one
  e->h_slot_of_row.assign(m, -1);
  e->h_free_slots.clear();
  e->h_pending_free.clear();
  for (int64_t s = e->kcap - 1; s >= 0; --s) e->h_free_slots.push_back((int32_t)s);
  e->colq_var = -1;
  if (any_structural_basic) {
    int64_t cnt = 0;
    for (int64_t r = 0; r < m; ++r) if (e->h_bvar[r] < ng) ++cnt;
    ST(ensure_lu_capacity(e, cnt));
    for (int64_t r = 0; r < m; ++r) {
      if (e->h_bvar[r] >= ng) continue;
      ST(fetch_column(e, e->h_bvar[r]));
      const int32_t slot = e->h_free_slots.back();
      e->h_free_slots.pop_back();
      if (!e->sparse)
        CU(cudaMemcpyAsync(e->Bcols + (size_t)slot * e->mld, e->colq, (size_t)m * 8, cudaMemcpyDeviceToDevice, e->stream));
      e->h_slot_of_row[r] = slot;
    }
  }
  e->initialized = true;
  ST(refactor_impl(e));
  ST(mark0(e));
  return MLP_OK;
}

mlp_status mlp_engine_set_primal_steepest_edge(mlp_engine* e, int32_t enable) {
  if (!e) return MLP_INVALID;
  e->enable_pse = enable;
  e->sel_valid = false;  // the score changes from d^2/gamma to |d| (solver.rs:713-716)
  return MLP_OK;
}

mlp_status mlp_refactor(mlp_engine* e, int64_t* lu_nnz) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(refactor_impl(e));
  ST(mark0(e));
  if (lu_nnz) *lu_nnz = e->lu_nnz;
  return MLP_OK;
}

mlp_status mlp_select_entering_primal(mlp_engine* e, mlp_entering* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int grid = std::min(cdiv(e->nt, 256), 1024);
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  if (!e->sel_valid)
    LAUNCH(e, k_select_primal, grid, 256, 0, e->d, e->gam, e->vflag, e->vpos, e->nt, e->n, e->c0, e->ng, e->enable_pse, l0.red_f,
           l0.red_i, l0.red_counter, e->xnb, e->d_res->flags, (Cand*)e->xsend);
  e->sel_valid = false;
  Cand w;
  w.var = -1;
  ST(exchange_candidates(e, &w));  // records s0_mark ahead of the run-ahead tail
  out->var = w.var;
  if (w.var < 0) { out->pos = -1; return MLP_OK; }
  out->pos = w.tie >> 32;
  out->score = w.key;
  out->obj_coeff = w.f[0];
  out->cur_val = w.f[1];
  return MLP_OK;
}

mlp_status mlp_ftran_col(mlp_engine* e, int64_t var) {
  if (!e || !e->initialized || var < 0 || var >= e->ng + e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  if (e->ftran_var == var) return MLP_OK;  // queued right behind the selection of `var`
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  e->ftran_var = -1;
  e->alpha_nnz_host = -1;
  ST(fetch_column(e, var));
  e->spec_var = -1;
  if (chain_fusable(e)) return chain_fused(e, var);
  ST(ftran(e, l0, e->colq, e->alpha, true));
  // |alpha|^2 (update_primal_sq_norms, 1136) and the stored size of col_coeffs (eta bookkeeping, 1096-1099)
  compact(e, l0, e->alpha, nullptr, nullptr, e->icnt + 1, e->scal + 2, e->touched_new);
  ST(mark0(e));
  if (e->enable_pse && e->overlap) ST(se_helper(e, var));  // runs ahead of the ratio test on lane 0
  return MLP_OK;
}

mlp_status mlp_ratio_primal(mlp_engine* e, int32_t sign, double max_step0, mlp_leaving* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int grid = std::min(cdiv(e->m, 256), 1024);
  Lane& l1 = e->lane[1];
  ST(begin1(e));
  LAUNCHS(e, l1.st, k_ratio_primal_1, grid, 256, 0, e->alpha, e->xB, e->loB, e->hiB, (int)e->m, sign, max_step0, l1.red_f,
          l1.red_counter, e->scal);
  LAUNCHS(e, l1.st, k_ratio_primal_2, grid, 256, 0, e->alpha, e->xB, e->loB, e->hiB, (int)e->m, sign, e->scal, l1.red_f, l1.red_i,
          l1.red_counter, l1.d_res);
  CU(cudaMemcpyAsync(&l1.d_res->i[1], e->icnt + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, l1.st));  // nnz(alpha_q)
  ST(mark1(e));
  ST(fetch_res(e, l1));
  e->alpha_nnz_host = (int64_t)(int32_t)(l1.h_res->i[1] & 0xffffffffLL);
  out->row = l1.h_res->i[0];
  out->coeff = l1.h_res->f[0];
  out->leaving_new_val = l1.h_res->f[1];
  out->basic_val = l1.h_res->f[2];
  out->ties = l1.h_res->i[2];
  out->near_ties = l1.h_res->i[3];
  if (out->row >= 0 && out->ties > 0) e->cnt.ratio_ties += 1;
  if (out->row >= 0 && out->near_ties > 0) e->cnt.ratio_near_ties += 1;
  return MLP_OK;
}

// row_dev (optional): the row is taken from device memory — the record k_select_row_dual left in d_res — so that the host
// can queue this BTRAN without having read the selection back (mlp_dual_select_ratio)
static mlp_status btran_unit_impl(mlp_engine* e, int64_t row, const long long* row_dev) {
  Lane& l1 = e->lane[1];
  ST(begin1(e));
  // c = e_row and, with etas, u = row `row` of E (solver.rs:1326-1330 for a unit vector) in one launch
  if (e->merge_small && e->K > 0 && e->K <= FE_MAXK) {  // short eta file: ... and s = (I+G)^-T u as well
    LAUNCHS(e, l1.st, k_unit_eta_t, cdiv(std::max<int64_t>(e->m, 32 * e->K), 256), 256, 0, e->work_mb, e->m, row, e->E, e->mld, e->Ginv, e->Kcap,
            (int)e->K, l1.tK2, row_dev);
    ST(btran(e, l1, e->work_mb, 0, e->rho, true, true));
  } else {
    LAUNCHS(e, l1.st, k_unit_and_gather, cdiv(std::max<int64_t>(e->m, e->K), 256), 256, 0, e->work_mb, e->m, row, e->E, e->mld, (int)e->K, l1.tK,
            row_dev);
    ST(btran(e, l1, e->work_mb, 0, e->rho, true));
  }
  // inv_basis_row_coeffs as a sparse list + |rho|^2 (solver.rs:683, 1160)
  compact(e, l1, e->rho, e->list_idx, e->list_val, e->icnt, e->scal + 1);
  ST(mark1(e));
  return MLP_OK;
}
mlp_status mlp_btran_unit(mlp_engine* e, int64_t row) {
  if (!e || !e->initialized || row < 0 || row >= e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  return btran_unit_impl(e, row, nullptr);
}

mlp_status mlp_price_row(mlp_engine* e) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(begin1(e));
  ST(price_list(e, e->lane[1], e->list_idx, e->list_val, e->icnt, 0, e->rho, e->rc, 0));
  return mark1(e);
}

mlp_status mlp_calc_row_coeffs(mlp_engine* e, int64_t row) {
  ST(mlp_btran_unit(e, row));
  return mlp_price_row(e);
}

static mlp_status select_row_dual_launch(mlp_engine* e) {
  const int grid = std::min(cdiv(e->m, 256), 1024);
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  LAUNCH(e, k_select_row_dual, grid, 256, 0, e->xB, e->loB, e->hiB, e->w, (int)e->m, e->enable_dse, l0.red_f, l0.red_i,
         l0.red_counter, e->d_res);
  return mark0(e);
}
mlp_status mlp_select_row_dual(mlp_engine* e, mlp_dual_row* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  Lane& l0 = e->lane[0];
  ST(select_row_dual_launch(e));
  ST(fetch_res(e, l0));
  out->row = e->h_res->i[0];
  out->val = e->h_res->f[0];
  out->min = e->h_res->f[1];
  out->max = e->h_res->f[2];
  e->dual_row_host = out->row;
  e->dual_row_val = out->val;
  return MLP_OK;
}

// choose_entering_col_dual (solver.rs:919-1021).  row_rec (optional): the leaving row's record in device memory (k_select_row_dual);
// leaving_diff_sign is then formed on the device and `lds` is ignored.
static mlp_status ratio_dual_impl(mlp_engine* e, int lds, const DevRes* row_rec, mlp_dual_entering* out) {
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  e->sel_valid = false;  // the candidate buffer is reused for the dual ratio test
  const int grid = std::min(cdiv(e->nt, 256), 1024);
  LAUNCH(e, k_ratio_dual_1, grid, 256, 0, e->rc, e->d, e->vflag, e->nt, lds, l0.red_f, l0.red_counter, e->scal, row_rec);
  if (e->world > 1) {  // Harris pass 1 is a min over ALL variables: all-gather the shard minima
    ST(e->comm->allgather(e->scal, e->xred, sizeof(double), e->stream));
    LAUNCH(e, k_min_small, 1, 1, 0, e->xred, e->world, e->scal);
  }
  LAUNCH(e, k_ratio_dual_2, grid, 256, 0, e->rc, e->d, e->vflag, e->vpos, e->xnb, e->nt, e->n, e->c0, e->ng, lds, e->scal,
         l0.red_f, l0.red_i, l0.red_counter, e->d_res->flags, (Cand*)e->xsend, e->rank == 0 ? 1 : 0, row_rec);
  Cand w;
  w.var = -1;
  ST(exchange_candidates(e, &w, true));
  out->var = w.var;
  if (w.var < 0) { out->pos = -1; return MLP_OK; }
  out->coeff = w.f[0];
  out->obj_coeff = w.f[1];
  out->cur_val = w.f[2];
  out->pos = (int64_t)w.f[3];
  out->ties = w.tie & 0xffff;
  out->near_ties = (w.tie >> 16) & 0xffff;
  if (out->ties > 0) e->cnt.ratio_ties += 1;
  if (out->near_ties > 0) e->cnt.ratio_near_ties += 1;
  return MLP_OK;
}
mlp_status mlp_ratio_dual(mlp_engine* e, int64_t row, double leaving_new_val, mlp_dual_entering* out) {
  if (!e || !e->initialized || row < 0 || row >= e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  // leaving_diff_sign = leaving_new_val > basic_var_vals[row] (solver.rs:925)
  double bv = e->dual_row_val;  // basic_var_vals[row] as choose_pivot_row_dual saw it; x_B has not changed since
  if (e->dual_row_host != row) ST(d2h(e, &bv, e->xB + row, sizeof(double)));
  return ratio_dual_impl(e, leaving_new_val > bv ? 1 : 0, nullptr, out);
}

// One host round trip for the first half of a dual iteration (solver.rs:529-531): choose_pivot_row_dual, calc_row_coeffs of
// the chosen row and choose_entering_col_dual are queued back to back — the row travels from kernel to kernel in device memory
// (k_select_row_dual's record: row, basic value, bounds) — and the host waits once, for the row's record and the winner's header
// together.  Same kernels, same arithmetic as mlp_select_row_dual + mlp_calc_row_coeffs + mlp_ratio_dual.  No infeasible row
// (row < 0): the queued BTRAN / price-out / ratio test run on a zero vector and find nothing; the caller stops on row < 0.
mlp_status mlp_dual_select_ratio(mlp_engine* e, mlp_dual_row* row_out, mlp_dual_entering* out) {
  if (!e || !e->initialized || !row_out || !out) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  Lane& l0 = e->lane[0];
  ST(select_row_dual_launch(e));
  CU(cudaMemcpyAsync(l0.h_res, l0.d_res, sizeof(DevRes), cudaMemcpyDeviceToHost, l0.st));  // read after the one wait below
  ST(mark0(e));
  e->cnt.d2h_bytes += (int64_t)sizeof(DevRes);
  ST(btran_unit_impl(e, -1, (const long long*)&e->d_res->i[0]));
  ST(mlp_price_row(e));
  e->dual_row_host = -1;
  ST(ratio_dual_impl(e, 0, e->d_res, out));  // waits for the winner's header: everything queued before it on lane 0 is done
  row_out->row = e->h_res->i[0];
  row_out->val = e->h_res->f[0];
  row_out->min = e->h_res->f[1];
  row_out->max = e->h_res->f[2];
  e->dual_row_host = row_out->row;
  e->dual_row_val = row_out->val;
  return MLP_OK;
}

mlp_status mlp_pivot(mlp_engine* e, const mlp_pivot_info* pi, mlp_pivot_result* out) {
  if (!e || !e->initialized || !pi || !out) return MLP_INVALID;
  if (pi->entering_var < 0 || pi->entering_var >= e->ng + e->m || pi->col < 0 || pi->col >= e->ng ||
      (pi->has_elem && (pi->row < 0 || pi->row >= e->m))) {
    set_err("pivot: entering_var / col / row out of range");
    return MLP_INVALID;
  }
  CU(cudaSetDevice(e->device));
  const int m = (int)e->m;
  Lane &l0 = e->lane[0], &l1 = e->lane[1];
  const int64_t q = pi->entering_var;
  const int64_t ql = to_local(e, q);
  out->leaving_var = -1;
  out->col_nnz = 0;
  out->refactored = 0;
  out->lu_nnz = e->lu_nnz;
  e->dual_row_host = -1;  // x_B is about to change
  if (!pi->has_elem) {  // solver.rs:1031-1042
    ST(begin0(e));
    e->sel_valid = false;
    LAUNCH(e, k_pivot_rows, cdiv(m, 256), 256, 0, e->alpha, e->tau, e->xB, e->w, m, -1, pi->entering_new_val, pi->entering_diff,
           1.0, 0, 0, e->scal, (double*)nullptr, e->d_res->flags, e->touched, e->touched_new);
    if (ql >= 0) LAUNCH(e, k_flip_var, 1, 1, 0, e->xnb, e->vflag, e->lo, e->hi, q, ql, pi->entering_new_val);
    ST(mark0(e));
    out->eta_count = e->K;
    return MLP_OK;
  }
  if (e->colq_var != q) { set_err("pivot: mlp_ftran_col(entering_var) must precede mlp_pivot"); return MLP_INVALID; }
  const int row = (int)pi->row;
  const int64_t lv = e->h_bvar[row];
  const int64_t lvl = to_local(e, lv);
  const double pivot_obj = pi->entering_obj_coeff / pi->coeff;  // solver.rs:1073
  bool do_refactor = pi->refactor != 0;
  if (!do_refactor && e->K >= e->Kcap) do_refactor = true;  // arena full
  if (do_refactor && e->refac_trace) { refac_stage(e, nullptr); e->refac_in_pivot = true; }
  // A refactorization between two true factorizations folds the eta file into C^-1 (refresh_inverse.cuh): then the eta of
  // THIS pivot is pushed like any other, so that the file describes the whole change of the basis.
  if (e->sparse) { e->h_eta_pos.push_back((int32_t)row); e->h_eta_leave.push_back(lv); }  // every basis change is on record
  e->K += 1;  // as can_refresh will see it
  bool refresh = do_refactor && e->K <= e->Kcap && e->pivots_since_lu + 1 < e->lu_every && can_refresh(e);
  e->K -= 1;
  // lane 1: tau = B^-1 rho (solver.rs:1157)
  ST(begin1(e));
  if (e->enable_dse) ST(ftran(e, l1, e->rho, e->tau));
  // lane 0: v = B^-T alpha_q and helper = N^T v, unless mlp_ftran_col already started them
  if (e->enable_pse) {
    if (e->spec_var != q) { ST(begin0(e)); ST(se_helper(e, q)); }
    if (e->overlap) CU(cudaStreamWaitEvent(l1.st, e->ev_vbtran, 0));  // the eta file is about to change
  }
  e->spec_var = -1;
  e->ftran_var = -1;
  // lane 1: row half of the pivot, eta push
  const bool push_eta = !do_refactor || refresh;
  double* eta_col = push_eta ? e->E + (size_t)e->K * e->mld : nullptr;
  LAUNCHS(e, l1.st, k_pivot_rows, cdiv(m, 256), 256, 0, e->alpha, e->tau, e->xB, e->w, m, row, pi->entering_new_val,
          pi->entering_diff, pi->coeff, 1, e->enable_dse, e->scal, eta_col, e->d_res->flags, e->touched, e->touched_new);
  e->pivots_since_lu += 1;
  if (push_eta) {
    const int prev = e->h_last_eta_of_row[row];
    const int K = (int)e->K;
    if (e->merge_small && K <= FE_MAXK)
      LAUNCHS(e, l1.st, k_eta_push, cdiv(K + 1, 8), 256, 0, e->E, e->mld, K, row, e->Ginv, e->Kcap, e->etaR, e->etaPrev, e->etaHead, e->etaLast, prev);
    else {
      LAUNCHS(e, l1.st, k_eta_grow, cdiv(std::max(K, 1), 256), 256, 0, e->E, e->mld, K, row, e->gK, e->etaR, e->etaPrev, e->etaHead, e->etaLast, prev);
      LAUNCHS(e, l1.st, k_eta_inv_row, cdiv(K + 1, 8), 256, 0, e->gK, e->Ginv, e->Kcap, K);
    }
    e->h_last_eta_of_row[row] = K;
    e->K += 1;
    e->cnt.etas_pushed += 1;
  }
  ST(mark1(e));
  // lane 0: variable half
  ST(begin0(e));
  {
    UpdSel a;
    a.partial = l0.partial; a.count_ptr = e->icnt + 2; a.lda = e->lda; a.slack_vals = e->vvec; a.helper = e->helper;
    a.finish = e->sparse ? 0 : 1;
    a.q = q; a.ql = ql; a.lvl = lvl; a.lv = lv; a.col = (int)pi->col; a.row = row;
    a.pivot_obj = pivot_obj; a.coeff = pi->coeff; a.leaving_new_val = pi->leaving_new_val; a.pse = e->enable_pse;
    LAUNCH(e, k_update_select, cdiv(e->nt, 256), 256, 0, a, e->d, e->gam, e->rc, e->xnb, e->vflag, e->vpos, e->bvar, e->loB,
           e->hiB, e->lo, e->hi, e->nt, e->n, e->m, e->c0, e->ng, e->scal, e->d_res->flags, l0.red_f, l0.red_i, l0.red_counter,
           e->d_res, (Cand*)e->xsend);
    e->sel_valid = true;  // the next choose_pivot scan is already done
  }
  // column cache: the leaving structural column frees its slot, the entering one takes a slot
  if (e->h_slot_of_row[row] >= 0) { e->h_pending_free.push_back(e->h_slot_of_row[row]); e->h_slot_of_row[row] = -1; }
  if (q < e->ng) {
    if (e->h_free_slots.empty()) {
      // doubles the cache; its content and slot numbers survive, the LU arrays do not: refactor in this pivot
      ST(ensure_lu_capacity(e, e->kcap + 1));
      do_refactor = true;
      refresh = false;
    }
    const int32_t slot = e->h_free_slots.back();
    e->h_free_slots.pop_back();
    if (!e->sparse)
      CU(cudaMemcpyAsync(e->Bcols + (size_t)slot * e->mld, e->colq, (size_t)m * 8, cudaMemcpyDeviceToDevice, e->stream));
    e->h_slot_of_row[row] = slot;
  }
  e->h_bvar[row] = q;
  const int par = (int)(e->pivot_seq & 1);
  e->pivot_seq += 1;
  if (!do_refactor && e->alpha_nnz_host >= 0 && e->async_pivot) {
    // Primal loop: everything the host needs is already known — the leaving variable from its mirror, nnz(alpha_q) from
    // the ratio test's read-back — so the call returns without waiting for the device; a non-finite norm surfaces with
    // the next selection (candidate header f[4]).  The host is then free to queue the next pivot's chain behind this one.
    ST(mark0(e));
    out->leaving_var = lv;
    out->col_nnz = e->alpha_nnz_host;
    e->alpha_nnz_host = -1;
    out->eta_count = e->K;
    return MLP_OK;
  }
  // one device->host read: status flags, leaving var, nnz(alpha)
  if (e->refac_in_pivot) refac_stage(e, "pivot: its own kernels (both lanes)");
  CU(cudaMemcpyAsync(&e->d_res->i[1], e->icnt + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
  ST(mark0(e));
  ST(fetch_res(e, l0));
  if (e->prof_on) ST(collect_profile(e, par));
  out->leaving_var = e->h_res->i[0];
  out->col_nnz = (int64_t)(int32_t)(e->h_res->i[1] & 0xffffffffLL);
  if (out->leaving_var != lv) { set_err("pivot: host/device basis mirrors diverged"); return MLP_INVALID; }
  if (e->h_res->flags[0] && e->world == 1) { set_err("non-finite steepest-edge norm"); return MLP_NONFINITE; }
  if (do_refactor) {
    ST(refactor_impl(e, refresh));
    ST(mark0(e));
    out->refactored = 1;
    out->lu_nnz = e->lu_nnz;
  }
  out->eta_count = e->K;
  return MLP_OK;
}

// ---------------------------------------------------------------------------------- incremental API
static mlp_status grow_rows(mlp_engine* e);
mlp_status mlp_get_var(mlp_engine* e, int64_t var, mlp_var_info* out) {
  if (!e || !e->initialized || !out || var < 0 || var >= e->ng + e->m) return MLP_INVALID;
  if (e->world != 1) { set_err("mlp_get_var: single-shard engines only"); return MLP_INVALID; }
  CU(cudaSetDevice(e->device));
  ST(begin0(e));
  const int64_t lv = to_local(e, var);
  uint8_t f = 0;
  int32_t pos = 0;
  ST(d2h(e, &f, e->vflag + lv, 1));
  ST(d2h(e, &pos, e->vpos + lv, 4));
  out->flags = f;
  out->pos_or_row = pos;
  if (f & MLP_BASIC) {
    out->obj_coeff = 0.0;
    ST(d2h(e, &out->value, e->xB + pos, 8));
  } else {
    ST(d2h(e, &out->obj_coeff, e->d + lv, 8));
    ST(d2h(e, &out->value, e->xnb + lv, 8));
  }
  return MLP_OK;
}
mlp_status mlp_set_nb_state(mlp_engine* e, int64_t var, uint32_t flags) {
  if (!e || !e->initialized || var < 0 || var >= e->ng + e->m) return MLP_INVALID;
  if (e->world != 1) { set_err("mlp_set_nb_state: single-shard engines only"); return MLP_INVALID; }
  CU(cudaSetDevice(e->device));
  ST(begin0(e));
  e->sel_valid = false;
  LAUNCH(e, k_set_var_state, 1, 1, 0, e->vflag, to_local(e, var), flags & (MLP_AT_MIN | MLP_AT_MAX | MLP_FIXED));
  return mark0(e);
}

mlp_status mlp_engine_add_row(mlp_engine* e, const double* coeffs, const double* slack_coeffs, double slack_min, double slack_max,
                              double rhs, mlp_add_row_result* out) {
  if (!e || !e->initialized || !coeffs || !out) return MLP_INVALID;
  if (e->world != 1) { set_err("add_row: single-shard engines only (row f2 is partly built)"); return MLP_INVALID; }
  CU(cudaSetDevice(e->device));
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  if (e->m >= e->mld) ST(grow_rows(e));
  Lane& l0 = e->lane[0];
  const int64_t m = e->m, n = e->n, r = m, lv = n + r, gv = e->ng + r;
  e->sel_valid = false;
  e->spec_var = e->ftran_var = -1;
  e->colq_var = -1;
  e->dual_row_host = -1;
  if (e->sparse) {
    // Sparse storage: the row is appended to the host CSR copy and the CSC copy, the segment table and the device
    // arrays are rebuilt from it — O(nnz), what the reference does per added constraint (solver.rs:598-610 rebuilds
    // its CSR row by row and calls to_csc).  The scalar work of 563-591 is done on the host in index order.
    std::vector<double> row(coeffs, coeffs + n);
    double rhs_new = rhs;
    if (slack_coeffs) {  // Gomory cut: substitute s_i = rhs_i - a_i x (see the dense path and DESIGN.md §8)
      std::vector<double> rhs_old((size_t)m), t((size_t)n, 0.0);
      ST(d2h(e, rhs_old.data(), e->rhs, m * 8));
      double dot = 0.0;
      for (int64_t i = 0; i < m; ++i) {
        const double g = slack_coeffs[i];
        if (g == 0.0) continue;
        for (int64_t q = e->h_csr_ptr[i]; q < e->h_csr_ptr[i + 1]; ++q) t[(size_t)e->h_csr_idx[q]] += g * e->h_csr_val[q];
        dot += g * rhs_old[(size_t)i];
      }
      for (int64_t j = 0; j < n; ++j) row[(size_t)j] -= t[(size_t)j];
      rhs_new = rhs - dot;
    }
    // basic value of the new slack: rhs - a . x over the structural variables (583-591), in index order
    std::vector<double> xnb((size_t)e->nt), xb((size_t)m);
    std::vector<uint8_t> fl((size_t)e->nt);
    std::vector<int32_t> pos((size_t)e->nt);
    ST(d2h(e, xnb.data(), e->xnb, e->nt * 8));
    ST(d2h(e, xb.data(), e->xB, m * 8));
    ST(d2h(e, fl.data(), e->vflag, e->nt));
    ST(d2h(e, pos.data(), e->vpos, e->nt * 4));
    double ax = 0.0;
    int64_t row_nnz = 0;
    for (int64_t j = 0; j < n; ++j) {
      const double c = row[(size_t)j];
      if (c == 0.0) continue;
      ++row_nnz;
      ax += c * ((fl[(size_t)j] & MLP_BASIC) ? xb[(size_t)pos[(size_t)j]] : xnb[(size_t)j]);
    }
    const double val = rhs_new - ax;
    for (int64_t j = 0; j < n; ++j)
      if (row[(size_t)j] != 0.0) { e->h_csr_idx.push_back((int32_t)j); e->h_csr_val.push_back(row[(size_t)j]); }
    e->h_csr_ptr.push_back(e->h_csr_ptr.back() + row_nnz);
    ST(sparse_upload(e, m + 1));
    double two[2] = {rhs_new, val};
    ST(h2d(e, e->scal + 7, two, sizeof(two)));  // [7] rhs of the new row, [8] its basic value
    LAUNCH(e, k_new_row_state, 1, 1, 0, r, lv, gv, slack_min, slack_max, e->scal + 8, e->scal + 7, e->lo, e->hi, e->cobj, e->d, e->gam,
           e->xnb, e->vflag, e->vpos, e->bvar, e->xB, e->loB, e->hiB, e->w, e->rhs, e->rowcover);
    CU(cudaStreamSynchronize(e->stream));
    e->m += 1;
    e->nt += 1;
    e->h_bvar.push_back(gv);
    e->h_slot_of_row.push_back(-1);
    e->h_last_eta_of_row.push_back(-1);
    ST(refactor_impl(e));  // basis_solver.reset (612)
    ST(mark0(e));
    if (e->enable_pse || e->enable_dse) {  // 615-630: the new tableau row extends the steepest-edge norms
      ST(mlp_calc_row_coeffs(e, r));
      ST(begin0(e));
      if (e->enable_pse) LAUNCH(e, k_add_sq, cdiv(e->nt, 256), 256, 0, e->gam, e->rc, e->vflag, e->nt);
      if (e->enable_dse) LAUNCH(e, k_copy1, 1, 1, 0, e->w + r, e->scal + 1);
      ST(mark0(e));
    }
    out->row = r;
    out->slack_var = gv;
    out->lu_nnz = e->lu_nnz;
    ST(d2h(e, &out->basic_val, e->xB + r, 8));
    ST(d2h(e, &out->rhs, e->rhs + r, 8));
    return MLP_OK;
  }
  double* rowA = e->A + r * e->lda;
  ST(h2d(e, rowA, coeffs, n * 8));
  double* d_rhs_new = e->scal + 7;
  if (slack_coeffs) {
    ST(h2d(e, e->work_m, slack_coeffs, m * 8));
    compact(e, l0, e->work_m, e->list_idx, e->list_val, e->icnt, e->scal + 8);
    if (e->price_tma)
      launch_price_tma(e, e->stream, e->list_idx, e->list_val, e->icnt, 0, l0.partial);
    else
      LAUNCH(e, k_price_partial<0>, price_grid(e), PR_THREADS, 0, e->A, e->lda, e->list_idx, e->list_val, e->icnt, 0, l0.partial);
    LAUNCH(e, k_row_combine, cdiv(n, 256), 256, 0, rowA, l0.partial, e->icnt, e->lda, n);
    LAUNCH(e, k_sub_dot, 1, 1024, 0, e->work_m, e->rhs, m, rhs, (const double*)nullptr, d_rhs_new);
  } else {
    ST(h2d(e, d_rhs_new, &rhs, 8));
  }
  // basic value of the new slack: rhs - a . x over the structural variables (583-591)
  LAUNCH(e, k_struct_values, cdiv(n, 256), 256, 0, e->xnb, e->xB, e->vflag, e->vpos, n, e->helper);
  LAUNCH(e, k_sub_dot, 1, 1024, 0, rowA, e->helper, n, 0.0, d_rhs_new, e->scal + 9);
  LAUNCH(e, k_new_row_state, 1, 1, 0, r, lv, gv, slack_min, slack_max, e->scal + 9, d_rhs_new, e->lo, e->hi, e->cobj, e->d, e->gam,
         e->xnb, e->vflag, e->vpos, e->bvar, e->xB, e->loB, e->hiB, e->w, e->rhs, e->rowcover);
  {  // cached basis columns: entry of the new row
    std::vector<int32_t> slots, vars;
    for (int64_t p = 0; p < m; ++p)
      if (e->h_slot_of_row[p] >= 0) { slots.push_back(e->h_slot_of_row[p]); vars.push_back((int32_t)e->h_bvar[p]); }
    if (!slots.empty()) {
      ST(h2d(e, e->vlist_idx, slots.data(), slots.size() * 4));
      ST(h2d(e, e->list_idx, vars.data(), vars.size() * 4));
      LAUNCH(e, k_bcols_new_row, cdiv((int64_t)slots.size(), 256), 256, 0, rowA, e->vlist_idx, e->list_idx, (int)slots.size(), e->mld,
             r, e->Bcols);
    }
    CU(cudaStreamSynchronize(e->stream));
  }
  e->m += 1;
  e->nt += 1;
  e->h_bvar.push_back(gv);
  e->h_slot_of_row.push_back(-1);
  e->h_last_eta_of_row.push_back(-1);
  ST(refactor_impl(e));  // basis_solver.reset (612)
  ST(mark0(e));
  if (e->enable_pse || e->enable_dse) {  // 615-630: the new tableau row extends the steepest-edge norms
    ST(mlp_calc_row_coeffs(e, r));
    ST(begin0(e));
    if (e->enable_pse) LAUNCH(e, k_add_sq, cdiv(e->nt, 256), 256, 0, e->gam, e->rc, e->vflag, e->nt);
    if (e->enable_dse) LAUNCH(e, k_copy1, 1, 1, 0, e->w + r, e->scal + 1);
    ST(mark0(e));
  }
  out->row = r;
  out->slack_var = gv;
  out->lu_nnz = e->lu_nnz;
  ST(d2h(e, &out->basic_val, e->xB + r, 8));
  ST(d2h(e, &out->rhs, e->rhs + r, 8));
  return MLP_OK;
}

// Solver: Clone (solver.rs:14, used by Solution: Clone lib.rs:313): device-to-device deep copy of the whole engine state,
// basis factors and eta file included, so the copy continues bit-identically.  new_mld > src->mld re-lays the row-indexed
// arrays out for a larger row capacity (grow_rows below); the leading dimension of Bcols / E changes with it.
static mlp_status clone_engine(mlp_engine* src, int64_t new_mld, mlp_engine** out) {
  if (!src || !out || !src->initialized) return MLP_INVALID;
  *out = nullptr;
  if (src->world != 1) { set_err("clone: single-shard engines only"); return MLP_INVALID; }
  if (new_mld < src->mld) return MLP_INVALID;
  CU(cudaSetDevice(src->device));
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(src->lane[l].st));
  mlp_engine* e = nullptr;
  ST(create_engine(src->device, src->m, src->ng, 0, 1, nullptr, &e, src->sparse, new_mld));
  mlp_status st = MLP_OK;
  auto cp = [&](void* dst, const void* from, size_t bytes) {
    if (st == MLP_OK && bytes && cudaMemcpyAsync(dst, from, bytes, cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess) {
      set_err("clone: device copy failed");
      st = MLP_CUDA_ERROR;
    }
  };
  // column-major block with leading dimension = row capacity: `cols` columns of src->mld rows each
  auto cp_cols = [&](double* dst, const double* from, size_t cols) {
    if (st != MLP_OK || !cols) return;
    if (cudaMemcpy2DAsync(dst, (size_t)e->mld * 8, from, (size_t)src->mld * 8, (size_t)src->mld * 8, cols, cudaMemcpyDeviceToDevice,
                          e->stream) != cudaSuccess) {
      set_err("clone: device copy failed");
      st = MLP_CUDA_ERROR;
    }
  };
  auto A = [&](mlp_status s2) { if (st == MLP_OK) st = s2; };
  const size_t ml = (size_t)src->mld, ntc = (size_t)src->n + ml, gt = (size_t)src->ng + ml;
  if (src->sparse) {  // matrix, CSC copy and segment table from the host CSR copy; then the core marks of the current factors
    e->h_csr_ptr = src->h_csr_ptr; e->h_csr_idx = src->h_csr_idx; e->h_csr_val = src->h_csr_val;
    A(dev_alloc(&e->corepos, (size_t)src->n)); A(dev_alloc(&e->rowcore, (size_t)e->mld));
    if (st == MLP_OK) A(sparse_upload(e, src->m));
    cp(e->corepos, src->corepos, (size_t)src->n * 4); cp(e->rowcore, src->rowcore, ml * 4);
  } else cp(e->A, src->A, ml * (size_t)src->lda * 8);
  cp(e->lo, src->lo, gt * 8); cp(e->hi, src->hi, gt * 8); cp(e->cobj, src->cobj, gt * 8);
  cp(e->d, src->d, ntc * 8); cp(e->gam, src->gam, ntc * 8); cp(e->xnb, src->xnb, ntc * 8);
  cp(e->vflag, src->vflag, ntc); cp(e->vpos, src->vpos, ntc * 4); cp(e->bvar, src->bvar, ml * 4);
  cp(e->xB, src->xB, ml * 8); cp(e->loB, src->loB, ml * 8); cp(e->hiB, src->hiB, ml * 8); cp(e->w, src->w, ml * 8);
  cp(e->rhs, src->rhs, ml * 8); cp(e->rowcover, src->rowcover, ml * 4);
  cp(e->alpha, src->alpha, ml * 8); cp(e->rho, src->rho, ml * 8); cp(e->rc, src->rc, ntc * 8); cp(e->helper, src->helper, ntc * 8);
  cp(e->scal, src->scal, 16 * 8); cp(e->icnt, src->icnt, 16 * 4);
  cp(e->d_res, src->d_res, sizeof(DevRes));
  if (st == MLP_OK && src->kcap > 0) {
    A(ensure_lu_capacity(e, src->kcap, true));
    if (st == MLP_OK && e->kcap != src->kcap) { set_err("clone: capacity mismatch"); st = MLP_INVALID; }
    const size_t kc = (size_t)src->kcap;
    cp(e->Jpos, src->Jpos, kc * 4); cp(e->Jslot, src->Jslot, kc * 4); cp(e->Rp, src->Rp, kc * 4);
    if (!src->sparse) cp_cols(e->Bcols, src->Bcols, kc);
    cp(e->LUc, src->LUc, kc * kc * 8); cp(e->Cinv, src->Cinv, kc * kc * 8);
    if (src->sparse) {
      cp(e->corevar, src->corevar, kc * 4); cp(e->cseg_first, src->cseg_first, (kc + 1) * 4);
      e->corevar_k = src->corevar_k;
      e->ncseg = src->ncseg;
      if (src->cseg_cap > 0) {
        e->cseg_cap = src->cseg_cap;
        A(dev_alloc(&e->cseg_id, (size_t)e->cseg_cap)); A(dev_alloc(&e->csum[0], (size_t)e->cseg_cap)); A(dev_alloc(&e->csum[1], (size_t)e->cseg_cap));
        cp(e->cseg_id, src->cseg_id, (size_t)src->ncseg * 4);
      }
    }
  }
  if (st == MLP_OK && src->Kcap > 0) {
    A(ensure_eta_capacity(e, src->Kcap, true));
    if (st == MLP_OK && e->Kcap != src->Kcap) { set_err("clone: capacity mismatch"); st = MLP_INVALID; }
    const size_t Kc = (size_t)src->Kcap;
    cp_cols(e->E, src->E, (size_t)src->K); cp(e->Ginv, src->Ginv, Kc * Kc * 8);
    cp(e->etaR, src->etaR, Kc * 4); cp(e->etaPrev, src->etaPrev, Kc * 4); cp(e->etaHead, src->etaHead, Kc * 4);
    cp(e->etaLast, src->etaLast, ml * 4);
  }
  cp(e->touched, src->touched, ml); cp(e->touched_new, src->touched_new, ml);
  if (st == MLP_OK && src->sparse && src->dcsr_ptr) {
    A(dev_alloc(&e->dcsr_ptr, (size_t)e->mld + 1)); A(dev_alloc(&e->dcsr_hist, (size_t)DCSR_CHUNKS * e->mld)); A(dev_alloc(&e->dcsr_cnt, (size_t)e->mld));
    e->dcsr_cap = src->dcsr_cap;
    if (e->dcsr_cap > 0) { A(dev_alloc(&e->dcsr_idx, (size_t)e->dcsr_cap)); A(dev_alloc(&e->dcsr_val, (size_t)e->dcsr_cap)); }
    cp(e->dcsr_ptr, src->dcsr_ptr, (ml + 1) * 8);
    if (e->dcsr_cap > 0) { cp(e->dcsr_idx, src->dcsr_idx, (size_t)e->dcsr_cap * 4); cp(e->dcsr_val, src->dcsr_val, (size_t)e->dcsr_cap * 8); }
  }
  if (st == MLP_OK && cudaStreamSynchronize(e->stream) != cudaSuccess) { set_err("clone: copy failed"); st = MLP_CUDA_ERROR; }
  if (st != MLP_OK) { destroy_engine(e); return st; }
  e->nt = src->nt;
  e->k = src->k; e->K = src->K; e->lu_nnz = src->lu_nnz;
  e->enable_pse = src->enable_pse; e->enable_dse = src->enable_dse;
  // the tuning state decides how reductions are tiled: the copy must round exactly like its source
  e->price_tma = src->price_tma; e->price_tile = src->price_tile; e->price_split = src->price_split;
  e->lane1_ldg = src->lane1_ldg; e->price_ctas = src->price_ctas; e->fused = src->fused; e->fused_max = src->fused_max;
  e->async_pivot = src->async_pivot; e->merge_small = src->merge_small;
  e->h_bvar = src->h_bvar; e->h_slot_of_row = src->h_slot_of_row; e->h_free_slots = src->h_free_slots;
  e->h_pending_free = src->h_pending_free; e->h_last_eta_of_row = src->h_last_eta_of_row;
  e->lu_every = src->lu_every; e->pivots_since_lu = src->pivots_since_lu; e->fill_true = src->fill_true; e->rf_tol = src->rf_tol;
  e->h_Jpos_f = src->h_Jpos_f; e->h_R_f = src->h_R_f; e->h_pos_core = src->h_pos_core; e->h_row_core = src->h_row_core;
  e->h_rowcover_f = src->h_rowcover_f; e->h_eta_pos = src->h_eta_pos; e->h_eta_leave = src->h_eta_leave;
  e->h_R_sorted = src->h_R_sorted; e->chg_complete = src->chg_complete && new_mld == src->mld;
  e->inv_valid = src->inv_valid && new_mld == src->mld;  // grow_rows refactorizes anyway
  e->cnt = src->cnt;
  e->initialized = true;
  ST(mark0(e));
  *out = e;
  return MLP_OK;
}
mlp_status mlp_engine_clone(mlp_engine* src, mlp_engine** out) { return src ? clone_engine(src, src->mld, out) : MLP_INVALID; }

// Row capacity exhausted (Solution::add_constraint / add_gomory_cut have no limit in the reference, lib.rs:368-423): double
// it.  Every row-indexed array and the leading dimension of Bcols / E depend on it, so the state is copied into a freshly
// laid-out engine (the clone path) whose guts then replace this handle's; the caller's pointer stays valid.
static mlp_status grow_rows(mlp_engine* e) {
  mlp_engine* bigger = nullptr;
  const int64_t want = e->mld + std::max<int64_t>(64, e->mld / 2);
  ST(clone_engine(e, want, &bigger));
  std::swap(*e, *bigger);
  destroy_engine(bigger);
  return MLP_OK;
}

// f4: recalc_basic_var_vals (solver.rs:1177-1197; dead code there, the TODO at 1024-1025 asks for it every ~1000
// pivots): x_B = B^-1 (rhs - N x_N) from scratch.  Off unless the caller asks (mlp_solver_set_recalc_period).
__global__ void k_masked_xnb(const double* __restrict__ xnb, const uint8_t* __restrict__ vflag, int64_t nt, double* __restrict__ out) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nt) out[v] = (vflag[v] & MLP_BASIC) ? 0.0 : xnb[v];
}
__global__ void k_sub_vec(double* __restrict__ y, const double* __restrict__ x, int64_t cnt) {
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) y[i] -= x[i];
}
mlp_status mlp_recalc_basic_vals(mlp_engine* e) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int64_t m = e->m, n = e->n, nt = e->nt;
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  e->spec_var = -1;
  e->dual_row_host = -1;
  if (e->K > 0) ST(refactor_impl(e));  // 1188-1191
  LAUNCH(e, k_masked_xnb, cdiv(nt, 256), 256, 0, e->xnb, e->vflag, nt, e->helper);
  double* part = e->world > 1 ? e->work_m : e->xred;
  if (e->sparse) LAUNCH(e, k_row_dot_csr, cdiv(m * 32, 256), 256, 0, e->csr_ptr, e->csr_idx, e->csr_val, m, e->c0, n, e->helper, part);
  else LAUNCH(e, k_row_dot, (unsigned)m, 256, 0, e->A, e->lda, n, e->helper, part);
  if (e->world > 1) ST(e->comm->allgather(part, e->xred, (size_t)m * sizeof(double), e->stream));
  LAUNCH(e, k_init_basic_vals, cdiv(m, 256), 256, 0, e->xred, e->world, (int)m, e->rhs, e->work_mb);  // rhs - A x_N (structural part)
  LAUNCH(e, k_sub_vec, cdiv(m, 256), 256, 0, e->work_mb, e->helper + n, m);                           // non-basic slacks: unit columns
  ST(ftran(e, l0, e->work_mb, e->xB));  // lu_factors.solve_dense (1193-1195): by constraint row in, by basis position out
  return mark0(e);
}

mlp_status mlp_recalc_obj_coeffs(mlp_engine* e, double* cur_obj_val) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int m = (int)e->m;
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  e->spec_var = -1;
  e->sel_valid = false;
  if (e->K > 0) ST(refactor_impl(e));  // solver.rs:1200-1203
  LAUNCH(e, k_gather_cB, cdiv(m, 256), 256, 0, e->cobj, e->bvar, m, e->work_m);
  ST(btran(e, l0, e->work_m, -1, e->vvec));  // multipliers y (1205-1214)
  compact(e, l0, e->vvec, e->vlist_idx, e->vlist_val, e->icnt + 2, e->scal + 3);
  ST(price_list(e, l0, e->vlist_idx, e->vlist_val, e->icnt + 2, 0, e->vvec, e->helper));
  LAUNCH(e, k_recalc_d, cdiv(e->nt, 256), 256, 0, e->cobj, e->helper, e->vflag, e->nt, e->n, e->c0, e->ng, e->d);
  LAUNCH(e, k_recalc_obj, 1, 1024, 0, e->cobj, e->bvar, e->xB, m, e->xnb, e->vflag, e->n, e->c0, e->ng, e->scal + 4);
  std::vector<double> parts((size_t)e->world, 0.0);
  double three[3];
  if (e->world > 1) {
    ST(e->comm->allgather(e->scal + 6, e->xred, sizeof(double), e->stream));
    ST(d2h(e, parts.data(), e->xred, sizeof(double) * e->world));
  }
  ST(d2h(e, three, e->scal + 4, sizeof(three)));
  if (e->world == 1) parts[0] = three[2];
  double tot = three[0];  // basic rows first (1225-1227), then the non-basic variables (1228-1230)
  tot += three[1];
  for (int r = 0; r < e->world; ++r) tot += parts[r];
  *cur_obj_val = tot;
  return mark0(e);
}

mlp_status mlp_download_f64(mlp_engine* e, int32_t which, double* out, int64_t count) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const double* src = nullptr;
  int64_t len = 0;
  switch (which) {
    case MLP_ARR_OBJ_COEFFS: src = e->d; len = e->nt; break;
    case MLP_ARR_PRIMAL_NORMS: src = e->gam; len = e->nt; break;
    case MLP_ARR_NB_VALS: src = e->xnb; len = e->nt; break;
    case MLP_ARR_BASIC_VALS: src = e->xB; len = e->m; break;
    case MLP_ARR_DUAL_NORMS: src = e->w; len = e->m; break;
    case MLP_ARR_COL_COEFFS: src = e->alpha; len = e->m; break;
    case MLP_ARR_INV_BASIS_ROW: src = e->rho; len = e->m; break;
    case MLP_ARR_ROW_COEFFS: src = e->rc; len = e->nt; break;
    case MLP_ARR_BASIC_MINS: src = e->loB; len = e->m; break;
    case MLP_ARR_BASIC_MAXS: src = e->hiB; len = e->m; break;
    case MLP_ARR_SE_HELPER: src = e->helper; len = e->nt; break;
    default: return MLP_INVALID;
  }
  if (count != len) { set_err("download: count mismatch"); return MLP_INVALID; }
  ST(begin0(e));
  return d2h(e, out, src, len * 8);
}
mlp_status mlp_download_basic_vars(mlp_engine* e, int64_t* out) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  std::vector<int32_t> tmp(e->m);
  ST(begin0(e));
  ST(d2h(e, tmp.data(), e->bvar, e->m * 4));
  for (int64_t i = 0; i < e->m; ++i) out[i] = tmp[i];
  return MLP_OK;
}
mlp_status mlp_download_var_state(mlp_engine* e, uint8_t* flags, int32_t* pos) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(begin0(e));
  ST(d2h(e, flags, e->vflag, e->nt));
  return d2h(e, pos, e->vpos, e->nt * 4);
}
mlp_status mlp_get_counters(mlp_engine* e, mlp_counters* out) {
  if (!e) return MLP_INVALID;
  *out = e->cnt;
  out->lu_nnz = e->lu_nnz;
  out->eta_count = e->K;
  return MLP_OK;
}
void* mlp_engine_stream(mlp_engine* e) { return e ? (void*)e->stream : nullptr; }
mlp_status mlp_engine_sync(mlp_engine* e) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  CU(cudaGetLastError());
  return MLP_OK;
}
mlp_status mlp_event_mark(mlp_engine* e, int32_t slot) {
  if (!e || slot < 0 || slot > 3) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  ST(begin0(e));
  CU(cudaEventRecord(e->ev[slot], e->stream));
  return MLP_OK;
}
mlp_status mlp_event_elapsed_ms(mlp_engine* e, int32_t a, int32_t b, double* ms) {
  if (!e || a < 0 || a > 3 || b < 0 || b > 3) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  CU(cudaEventSynchronize(e->ev[b]));
  float f = 0.f;
  CU(cudaEventElapsedTime(&f, e->ev[a], e->ev[b]));
  *ms = f;
  return MLP_OK;
}
mlp_status mlp_profile_enable(mlp_engine* e, int32_t on) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  if (e->prof_on) for (int par = 0; par < 2; ++par) ST(collect_profile(e, par));
  e->prof_on = on;
  std::memset(e->ppending, 0, sizeof(e->ppending));
  if (on) e->prof = mlp_profile{};
  return MLP_OK;
}
mlp_status mlp_profile_get(mlp_engine* e, mlp_profile* out) {
  if (!e || !out) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  for (int par = 0; par < 2; ++par) ST(collect_profile(e, par));
  *out = e->prof;
  return MLP_OK;
}

mlp_status mlp_engine_set_tuning(mlp_engine* e, int32_t knob, int32_t value) {
  if (!e) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  switch (knob) {
    case MLP_TUNE_PRICE_TILE:
      if (value == 0) { choose_price_tiling(e->lda, e->m, e->sm_count, &e->price_tile, &e->price_split); break; }  // automatic
      if (value < 128 || value > 4096 || value % 64 != 0) { set_err("set_tuning: tile width must be a multiple of 64 in [128, 4096]"); return MLP_INVALID; }
      e->price_tile = value;
      e->price_split = 1;
      break;
    case MLP_TUNE_PRICE_SPLIT:
      if ((value != 1 && value != 2 && value != 4) || e->price_tile / value < 128 || (e->price_tile / value) % 16 != 0) {
        set_err("set_tuning: split must be 1, 2 or 4 and leave slices of >= 128 columns");
        return MLP_INVALID;
      }
      e->price_split = value;
      break;
    case MLP_TUNE_LANE1_LDG: e->lane1_ldg = value != 0; break;
    case MLP_TUNE_FUSED: {
      int nb = 0;
      if (value && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_chain_primal, FZ_T, 0) != cudaSuccess || nb < 1)) {
        cudaGetLastError();
        set_err("set_tuning: the fused chain cannot be launched cooperatively on this device");
        return MLP_INVALID;
      }
      e->fused = value != 0;
      break;
    }
    case MLP_TUNE_FUSED_MAX: e->fused_max = std::max(0, std::min(FZ_MAX, (int)value)); break;
    case MLP_TUNE_LU_EVERY: e->lu_every = std::max(0, (int)value); break;
    default: set_err("set_tuning: unknown knob"); return MLP_INVALID;
  }
  e->spec_var = -1;
  return MLP_OK;
}

mlp_status mlp_engine_get_tuning(mlp_engine* e, int32_t knob, int32_t* value) {
  if (!e || !value) return MLP_INVALID;
  switch (knob) {
    case MLP_TUNE_PRICE_TILE: *value = e->price_tile; break;
    case MLP_TUNE_PRICE_SPLIT: *value = e->price_split; break;
    case MLP_TUNE_LANE1_LDG: *value = e->lane1_ldg; break;
    case MLP_TUNE_FUSED: *value = e->fused; break;
    case MLP_TUNE_FUSED_MAX: *value = e->fused_max; break;
    case MLP_TUNE_LU_EVERY: *value = (int32_t)e->lu_every; break;
    default: set_err("get_tuning: unknown knob"); return MLP_INVALID;
  }
  return MLP_OK;
}

mlp_status mlp_bench_price_dense(mlp_engine* e, int32_t iters, double* ms_per_launch, int64_t* bytes_per_launch) {
  if (!e || iters <= 0) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int m = (int)e->m;
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  e->spec_var = -1;
  LAUNCH(e, k_fill, cdiv(m, 256), 256, 0, e->vvec, (int64_t)m, 0.5);
  compact(e, l0, e->vvec, e->vlist_idx, e->vlist_val, e->icnt + 2, e->scal + 3);
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  ST(price_list(e, l0, e->vlist_idx, e->vlist_val, e->icnt + 2, 0, e->vvec, e->helper));  // warm-up
  CU(cudaEventRecord(a, e->stream));
  for (int i = 0; i < iters; ++i) ST(price_list(e, l0, e->vlist_idx, e->vlist_val, e->icnt + 2, 0, e->vvec, e->helper));
  CU(cudaEventRecord(b, e->stream));
  CU(cudaEventSynchronize(b));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *ms_per_launch = (double)ms / iters;
  // algorithmic bytes: 8 n s + 8 s + 8 n  with s = m, n = this shard's columns
  *bytes_per_launch = e->sparse ? 12 * e->nnz_loc + 8 * (int64_t)m + 8 * e->nt : 8 * e->n * (int64_t)m + 8 * (int64_t)m + 8 * e->n;
  return MLP_OK;
}

}  // extern "C"
