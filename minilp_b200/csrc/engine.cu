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

// ------------------------------------------------------------------------------------------------ K5 price-out
// calc_row_coeffs price-out (solver.rs:685-692), the N^T v product of update_primal_sq_norms (1117-1132), the full
// c_N - N^T y of recalc_obj_coeffs (1216-1222) and the column norms of try_new (297-299).  Row-gather GEMV^T over
// row-major A: a CTA owns a 512-column tile and one chunk of the multiplier's support; each thread accumulates two
// adjacent columns with 128-bit streaming loads, 8 rows in flight; rows of a chunk are taken in list order;
// (row, weight) pairs are staged through shared memory.  The chunking depends ONLY on the support size s
// (C = clamp(ceil(s/128), 1, 64)), never on the grid or the shard width, so a column's sum is bit-identical however
// the columns are sharded.  Chunk partials are reduced in chunk order by k_price_finish — no atomics.
constexpr int PR_THREADS = 256;
constexpr int PR_TILE = PR_THREADS * 2;
constexpr int PR_BATCH = 256;
constexpr int PR_UNROLL = 8;
constexpr int PR_MAXC = 64;
__host__ __device__ __forceinline__ int price_chunks_for(int s) {
  int c = (s + 127) / 128;
  return c < 1 ? 1 : (c > PR_MAXC ? PR_MAXC : c);
}

template <int MODE>  // 0: sum_r w_r * A[r,j]   1: sum_r A[r,j]^2
__global__ void __launch_bounds__(PR_THREADS)
k_price_partial(const double* __restrict__ A, int64_t lda, const int32_t* __restrict__ rows,
                const double* __restrict__ wts, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                double* __restrict__ partial) {
  pdl_wait();
  __shared__ int32_t srow[PR_BATCH];
  __shared__ double sw[PR_BATCH];
  const int s = count_ptr ? *count_ptr : fixed_count;
  const int C = price_chunks_for(s);
  const int L = (s + C - 1) / C;
  const int tiles = (int)((lda + PR_TILE - 1) / PR_TILE);
  // Persistent CTAs: the grid is sized to a fixed number of CTAs per SM (not to the work), so that the rest of each SM stays
  // free for the latency-bound kernels of the other lane; work item = (column tile, support chunk).
  for (int item = blockIdx.x; item < tiles * C; item += gridDim.x) {
    const int tile = item % tiles, chunk = item / tiles;
    const int k0 = chunk * L;
    const int k1 = min(s, k0 + L);
    const int64_t col = ((int64_t)tile * PR_THREADS + threadIdx.x) * 2;
    const bool active = col < lda;
    double acc0 = 0.0, acc1 = 0.0;
    const double* base = A + col;
    for (int kb = k0; kb < k1; kb += PR_BATCH) {
      const int nb = min(PR_BATCH, k1 - kb);
      __syncthreads();
      for (int t = threadIdx.x; t < nb; t += PR_THREADS) {
        srow[t] = rows ? rows[kb + t] : kb + t;
        sw[t] = (MODE == 0) ? wts[kb + t] : 1.0;
      }
      __syncthreads();
      if (active) {
        int i = 0;
        for (; i + PR_UNROLL <= nb; i += PR_UNROLL) {
          double2 v[PR_UNROLL];
#pragma unroll
          for (int u = 0; u < PR_UNROLL; ++u)
            v[u] = __ldcs(reinterpret_cast<const double2*>(base + (int64_t)srow[i + u] * lda));
#pragma unroll
          for (int u = 0; u < PR_UNROLL; ++u) {
            if (MODE == 0) {
              const double wv = sw[i + u];
              acc0 += wv * v[u].x;
              acc1 += wv * v[u].y;
            } else {
              acc0 += v[u].x * v[u].x;
              acc1 += v[u].y * v[u].y;
            }
          }
        }
        for (; i < nb; ++i) {
          const double2 v = __ldcs(reinterpret_cast<const double2*>(base + (int64_t)srow[i] * lda));
          if (MODE == 0) {
            const double wv = sw[i];
            acc0 += wv * v.x;
            acc1 += wv * v.y;
          } else {
            acc0 += v.x * v.x;
            acc1 += v.y * v.y;
          }
        }
      }
    }
    if (active) {
      double2 o;
      o.x = acc0;
      o.y = acc1;
      *reinterpret_cast<double2*>(partial + (int64_t)chunk * lda + col) = o;
    }
  }
}

// ---- bulk-copy (TMA) form of the same price-out --------------------------------------------------------------
// The LDG kernel above needs 6 resident CTAs per SM (all registers) to keep enough bytes in flight; that starves the
// latency-bound kernels of lane 1 that should run beside it.  Here the bytes in flight live in SHARED memory instead:
// one CTA per SM, a producer warp gathers the listed rows with 1-D bulk copies (cp.async.bulk, 4 KB row segments,
// L2 evict-first) into a TP_STAGES-deep ring guarded by mbarriers, eight consumer warps accumulate.  Work items,
// chunking and the per-column accumulation order (list order within a chunk, thread t owns columns 2t, 2t+1 of the
// tile) are those of k_price_partial, so the partial sums are bit-identical.
constexpr int TP_STAGE_BYTES = 32768;             // one stage: R rows x tile_cols x 8 B, R = 32768 / (tile_cols * 8) <= 32
constexpr int TP_STAGES = 6;                      // ring depth: 6 x 32 KB = 192 KB in flight per SM
constexpr int TP_MAXROWS = 32;                    // one row per producer lane
constexpr int TP_CONSUMERS = PR_THREADS;          // 8 warps
constexpr int TP_THREADS = TP_CONSUMERS + 32;     // + producer warp
constexpr size_t TP_SMEM = (size_t)TP_STAGES * TP_STAGE_BYTES + TP_STAGES * TP_MAXROWS * 8 + 2 * TP_STAGES * 8 + 16;
// Tile width: 512 columns (4 KB row segments).  Narrower tiles were measured and lose: 6.77 TB/s at 512, 6.19 at 256,
// 4.30 at 128 columns (more, smaller bulk copies per byte); a narrow column block of a sharded engine has fewer work
// items per SM, but the tail still has enough SMs active to saturate HBM.  MLP_PRICE_TILE overrides for experiments.
// Choice of the tiling (host), from the measured sweeps in profiles/r01d_price_sweep.md (B200, isolated dense N^T v,
// m = 50k; GB/s at 512-column tiles -> at the width chosen here):  n_loc 50 000: 6760 -> 7068 (1280);  25 000: 6735 -> 7070
// (1280);  12 500: 6668 -> 7102 (4096);  6 250: 6125 -> 6998 (4096).  Wider row segments mean fewer, larger bulk copies and
// longer contiguous DRAM bursts, and that matters more the shorter the rows of the local block are; widths whose rows fill
// a 32 KB stage badly (2560 columns = 20 KB: one row per stage) lose the bytes in flight again.  The kernel is HBM-bound
// with ~28 MB in flight, so a partly filled last round of work items costs little (a round model that predicted gains
// from balancing it did not survive the measurement; the tail split stays available as a knob, default 1).
static void choose_price_tiling(int64_t lda, int64_t /*m*/, int /*G*/, int* tile, int* split) {
  *split = 1;
  const int w = lda < 20000 ? 4096 : 1280;
  const int need = (int)std::min<int64_t>(4096, (lda + 63) / 64 * 64);  // never wider than the block itself
  *tile = std::max(128, std::min(w, need));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)  // L2 evict-first: A is streamed once per pivot
      : "memory");
}

// NV = column pairs per consumer thread: a tile is up to NV * 512 columns wide (NV * 4 KB row segments).  Thread t owns
// column pairs t, t + 256, ... of the tile; every column still accumulates its rows in list order, so the partial sums
// do not depend on the tiling.
//
// Work items and load balance.  An item is (column tile, support chunk); CTA b takes items b, b + G, ... (G CTAs).  With
// tiles * C items the last round is only partly filled — at 50k columns and 1024-column tiles 3136 items are 21.2
// rounds, i.e. 4 % of the kernel runs with 80 % of the SMs idle.  So only the items of the FULL rounds keep the whole tile
// width; the items of the last, partial round are cut into `split` column slices each, which spreads that round over
// all CTAs again (narrower row segments are less efficient, but only the tail pays that).
struct PriceItem {
  int chunk;
  int64_t col0;
  int cols;   // columns actually present (0: the slice lies beyond the matrix)
  int width;  // nominal width: row stride of the stage in shared memory
};
__device__ __forceinline__ PriceItem price_item(int idx, int main_items, int tiles, int tile_cols, int split, int64_t lda) {
  PriceItem it;
  int item, sub = 0;
  it.width = tile_cols;
  if (idx < main_items) item = idx;
  else {
    const int j = idx - main_items;
    item = main_items + j / split;
    sub = j % split;
    it.width = tile_cols / split;
  }
  it.chunk = item / tiles;
  it.col0 = (int64_t)(item % tiles) * tile_cols + (int64_t)sub * it.width;
  const int64_t tile_end = min(lda, (int64_t)(item % tiles + 1) * tile_cols);
  const int64_t c = min((int64_t)it.width, tile_end - it.col0);
  it.cols = c > 0 ? (int)c : 0;
  return it;
}
template <int NV>
__global__ void __launch_bounds__(TP_THREADS, 1)
k_price_partial_tma(const double* __restrict__ A, int64_t lda, const int32_t* __restrict__ rows,
                    const double* __restrict__ wts, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                    double* __restrict__ partial, int tile_cols, int split) {
  pdl_wait();
  extern __shared__ __align__(128) unsigned char tp_smem[];
  double* sw = reinterpret_cast<double*>(tp_smem + (size_t)TP_STAGES * TP_STAGE_BYTES);   // [stage][row] weights
  uint64_t* full = reinterpret_cast<uint64_t*>(sw + TP_STAGES * TP_MAXROWS);
  uint64_t* empty = full + TP_STAGES;
  const int s = count_ptr ? *count_ptr : fixed_count;
  const int C = price_chunks_for(s);
  const int L = (s + C - 1) / C;
  const int tiles = (int)((lda + tile_cols - 1) / tile_cols);
  const int items = tiles * C;
  const int main_items = items / (int)gridDim.x * (int)gridDim.x;
  const int total = main_items + (items - main_items) * split;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int q = 0; q < TP_STAGES; ++q) { mbar_init(full + q, 1); mbar_init(empty + q, TP_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  uint32_t it = 0;  // stages handled so far by this role: slot = it % TP_STAGES, phase = (it / TP_STAGES) & 1
  if (warp == TP_CONSUMERS / 32) {
    // ---------------- producer warp: lane r < R fetches row r of the stage
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      const PriceItem pi = price_item(idx, main_items, tiles, tile_cols, split, lda);
      if (pi.cols == 0) continue;
      const int k0 = pi.chunk * L, k1 = min(s, k0 + L);
      const int tile_bytes = pi.width * 8;
      const int R = min(TP_MAXROWS, TP_STAGE_BYTES / tile_bytes);  // rows per stage
      const uint32_t tbytes = (uint32_t)pi.cols * 8u;
      for (int kb = k0; kb < k1; kb += R, ++it) {
        const int nr = min(R, k1 - kb);
        const int slot = it % TP_STAGES;
        int32_t r = 0;
        double wv = 0.0;
        if (lane < nr) { r = rows[kb + lane]; wv = wts[kb + lane]; }  // issued before the wait: latency overlaps
        mbar_wait(empty + slot, ((it / TP_STAGES) & 1) ^ 1);
        if (lane < nr) sw[slot * TP_MAXROWS + lane] = wv;
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(full + slot, tbytes * nr);
        __syncwarp();
        if (lane < nr)
          bulk_g2s(tp_smem + (size_t)slot * TP_STAGE_BYTES + (size_t)lane * tile_bytes, A + (int64_t)r * lda + pi.col0, tbytes,
                   full + slot);
      }
    }
  } else {
    // ---------------- consumers: thread t owns column pairs t + 256 v (v < NV) of the tile
    const int t = threadIdx.x;
    constexpr int RU = 8 / NV;            // rows per unrolled batch: 8 loads in flight per thread
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      const PriceItem pi = price_item(idx, main_items, tiles, tile_cols, split, lda);
      if (pi.cols == 0) continue;
      const int k0 = pi.chunk * L, k1 = min(s, k0 + L);
      const int tile_bytes = pi.width * 8;
      const int R = min(TP_MAXROWS, TP_STAGE_BYTES / tile_bytes);
      const int rstride = tile_bytes / 16;  // double2 per row
      const int64_t col = pi.col0 + 2 * t;
      bool active[NV];
      double acc0[NV], acc1[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        active[v] = 2 * (t + TP_CONSUMERS * v) < pi.cols;
        acc0[v] = 0.0;
        acc1[v] = 0.0;
      }
      for (int kb = k0; kb < k1; kb += R, ++it) {
        const int nr = min(R, k1 - kb);
        const int slot = it % TP_STAGES;
        mbar_wait(full + slot, (it / TP_STAGES) & 1);
        if (active[0]) {
          const double2* p = reinterpret_cast<const double2*>(tp_smem + (size_t)slot * TP_STAGE_BYTES) + t;
          const double* w = sw + slot * TP_MAXROWS;
          int u0 = 0;
          for (; u0 + RU <= nr; u0 += RU) {
            double2 x[RU][NV];
#pragma unroll
            for (int u = 0; u < RU; ++u)
#pragma unroll
              for (int v = 0; v < NV; ++v)
                if (NV == 1 || active[v]) x[u][v] = p[(u0 + u) * rstride + TP_CONSUMERS * v];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
              const double wv = w[u0 + u];
#pragma unroll
              for (int v = 0; v < NV; ++v)
                if (NV == 1 || active[v]) {
                  acc0[v] += wv * x[u][v].x;
                  acc1[v] += wv * x[u][v].y;
                }
            }
          }
          for (; u0 < nr; ++u0) {
            const double wv = w[u0];
#pragma unroll
            for (int v = 0; v < NV; ++v)
              if (NV == 1 || active[v]) {
                const double2 x = p[u0 * rstride + TP_CONSUMERS * v];
                acc0[v] += wv * x.x;
                acc1[v] += wv * x.y;
              }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (active[v]) {
          double2 o;
          o.x = acc0[v];
          o.y = acc1[v];
          *reinterpret_cast<double2*>(partial + (int64_t)pi.chunk * lda + col + 2 * TP_CONSUMERS * v) = o;
        }
    }
  }
}

// Reduce chunk partials in chunk order; slack columns of [A|I] contribute rho_i (the `I` part of the CSR row,
// solver.rs:250); basic variables are not part of row_coeffs (solver.rs:688).
// mode 0: out = sum   mode 1: out = sum + 1 (primal edge norms, solver.rs:298)
__global__ void k_price_finish(const double* __restrict__ partial, const int32_t* __restrict__ count_ptr, int32_t fixed_count,
                               int64_t lda, int64_t n, int64_t m, const double* __restrict__ slack_vals,
                               const uint8_t* __restrict__ vflag, double* __restrict__ out, int mode) {
  pdl_wait();
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n + m) return;
  const int C = price_chunks_for(count_ptr ? *count_ptr : fixed_count);
  double r;
  if (v < n) {
    r = 0.0;
    for (int c = 0; c < C; ++c) r += partial[(int64_t)c * lda + v];
    if (mode == 1) r += 1.0;
  } else {
    r = (mode == 1) ? 2.0 : slack_vals[v - n];  // |e_i|^2 + 1
  }
  if (mode == 0 && (vflag[v] & MLP_BASIC)) r = 0.0;
  out[v] = r;
}

// ------------------------------------------------------------------------------------------------ columns
// rhs.set(column of var) (solver.rs:672-675, sparse.rs:103): column of [A|I] of LOCAL variable lv as a dense m-vector
__global__ void k_load_col(const double* __restrict__ A, int64_t lda, int64_t n, int m, int64_t lv, double* __restrict__ dst) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  dst[i] = lv < n ? A[(int64_t)i * lda + lv] : ((int64_t)i == lv - n ? 1.0 : 0.0);
}
// same, for the variable named by a candidate header that is still on the device (no host round trip)
__global__ void k_cand_load_col(const double* __restrict__ A, int64_t lda, int64_t n, int64_t c0, int64_t ng, int m,
                                const Cand* __restrict__ cand, double* __restrict__ dst, Cand* __restrict__ win_out) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && win_out) *win_out = *cand;  // single shard: the candidate IS the winner
  if (i >= m) return;
  const long long g = cand->var;
  if (g < 0) { dst[i] = 0.0; return; }
  const int64_t lv = g >= ng ? n + (g - ng) : g - c0;
  dst[i] = lv < n ? A[(int64_t)i * lda + lv] : ((int64_t)i == lv - n ? 1.0 : 0.0);
}

// FTRAN tail: alpha[pos] for slack positions = a_i - (D1 x)_i ; alpha[Jpos[t]] = x[t]   (U-solve of the
// identity-bordered basis, lu.rs:93 with B = [D | E_S]).  The t-th dense column lives in cache slot Jslot[t].
__global__ void __launch_bounds__(256) k_ftran_finish(const double* __restrict__ Bcols, int64_t ldb, int m, int k,
                                                       const double* __restrict__ xk, const double* __restrict__ rhs0,
                                                       const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                                       const int32_t* __restrict__ Jslot, double* __restrict__ out,
                                                       const uint8_t* __restrict__ touched, uint8_t* __restrict__ touched_new) {
  pdl_wait();
  __shared__ double ts[512];
  __shared__ int32_t sl[512];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = i < m ? rhs0[i] : 0.0;
  const int cov = i < m ? rowcover[i] : -1;
  for (int j0 = 0; j0 < k; j0 += 512) {
    const int nj = min(512, k - j0);
    __syncthreads();
    for (int q = threadIdx.x; q < nj; q += blockDim.x) { ts[q] = xk[j0 + q]; sl[q] = Jslot[j0 + q]; }
    __syncthreads();
    if (cov >= 0) {
      const double* p = Bcols + i;
      int j = 0;
      for (; j + 16 <= nj; j += 16) {  // 16 loads in flight, subtraction still in column order
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = p[(int64_t)sl[j + u] * ldb];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j + 4 <= nj; j += 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[(int64_t)sl[j + u] * ldb];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc -= ts[j + u] * v[u];
      }
      for (; j < nj; ++j) acc -= ts[j] * p[(int64_t)sl[j] * ldb];
    }
  }
  if (cov >= 0) { out[cov] = acc; if (touched_new) touched_new[cov] = (uint8_t)((acc != 0.0) | (touched[cov] != 0)); }
  if (i < k) {
    const int p = Jpos[i];
    const double xv = xk[i];
    out[p] = xv;
    if (touched_new) touched_new[p] = (uint8_t)((xv != 0.0) | (touched[p] != 0));  // k_touch_mark, folded in
  }
}
// FTRAN tail after a column-group split (k_tall_part): alpha_slack = a_S - sum_g part[g], alpha[Jpos[t]] = x[t]
__global__ void k_ftran_finish_parts(const double* __restrict__ part, int G, int64_t pld, int m, int k,
                                     const double* __restrict__ xk, const double* __restrict__ rhs0,
                                     const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                     double* __restrict__ out, const uint8_t* __restrict__ touched,
                                     uint8_t* __restrict__ touched_new) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    const int cov = rowcover[i];
    if (cov >= 0) {
      double tsum = 0.0;
      for (int g = 0; g < G; ++g) tsum += part[(int64_t)g * pld + i];
      const double v = rhs0[i] - tsum;
      out[cov] = v;
      if (touched_new) touched_new[cov] = (uint8_t)((v != 0.0) | (touched[cov] != 0));
    }
  }
  if (i < k) {
    const int p = Jpos[i];
    const double xv = xk[i];
    out[p] = xv;
    if (touched_new) touched_new[p] = (uint8_t)((xv != 0.0) | (touched[p] != 0));
  }
}
// core C = D[R,:] (k x k, column-major) from the column cache
__global__ void k_extract_core(const double* __restrict__ Bcols, int64_t ldb, int k, const int32_t* __restrict__ Rp,
                               const int32_t* __restrict__ Jslot, double* __restrict__ C, int64_t ld) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (i < k && t < k) C[(int64_t)t * ld + i] = Bcols[(int64_t)Jslot[t] * ldb + Rp[i]];
}
// BTRAN: right-hand side of the core solve, rhs_t = c[Jpos[t]] - sum_i Bcols[i, slot_t] cov_i; CTA (t, s) reduces a row
// slice, k_gemv_t_fin adds the slices in order.
__global__ void __launch_bounds__(256) k_core_rhs_part(const double* __restrict__ Bcols, int64_t ldb, int rows, int k,
                                                        const int32_t* __restrict__ Jslot, const double* __restrict__ x,
                                                        double* __restrict__ part) {
  pdl_wait();
  __shared__ double sm[32];
  const int j = blockIdx.x, S = gridDim.y, sidx = blockIdx.y;
  const int L = (rows + S - 1) / S;
  const int r0 = sidx * L, r1 = min(rows, r0 + L);
  const double* p = Bcols + (int64_t)Jslot[j] * ldb;
  double acc = 0.0;
  int i = r0 + threadIdx.x;
  const int st = blockDim.x;
  for (; i + 7 * st < r1; i += 8 * st) {  // 8 row pairs in flight per thread, added in row order
    double a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = p[i + u * st]; b[u] = x[i + u * st]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += a[u] * b[u];
  }
  for (; i < r1; i += st) acc += p[i] * x[i];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) part[(int64_t)sidx * k + j] = tot;
}

// ------------------------------------------------------------------------------------------------ sparse storage
// Sparse A (CSR + CSC, u32 indices).  The basis-inverse machinery is shared with the dense engine: basis columns are
// expanded into the dense column cache when they enter, so only three things read the sparse matrix — the column
// load, the price-out and the set-up passes.
// rhs.set(column) (solver.rs:672-675): dst is zero-filled by the caller; var < 0 comes from a candidate header.
// Every shard of a sparse-storage engine holds the WHOLE matrix (12 nnz bytes: small next to HBM; the basis operations
// need the basic columns wherever they price): n and lv are GLOBAL here.
__global__ void k_load_col_csc(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
                               int64_t n, int64_t lv_arg, const Cand* __restrict__ cand, double* __restrict__ dst,
                               Cand* __restrict__ win_out) {
  pdl_wait();
  int64_t lv = lv_arg;
  if (cand && win_out && blockIdx.x == 0 && threadIdx.x == 0) *win_out = *cand;
  if (cand) {
    if (cand->var < 0) return;
    lv = cand->var;  // GLOBAL variable index
  }
  if (lv >= n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[lv - n] = 1.0;
    return;
  }
  const int64_t b = ptr[lv], e = ptr[lv + 1];
  for (int64_t t = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < e; t += (int64_t)gridDim.x * blockDim.x) dst[idx[t]] = val[t];
}
// Price-out over the CSC copy (calc_row_coeffs 685-692, update_primal_sq_norms 1117-1132, recalc_obj_coeffs 1216-1222,
// column norms 297-299): gathers the DENSE multiplier vector w at each column's row indices — 12 bytes per stored entry.
// Column lengths are power-law distributed (one column of a netlib-like LP can hold 10^5 entries), so the unit of work is
// a SEGMENT of at most CSC_SEG consecutive entries of one column (table built once at creation): pass 1, one warp per
// segment, rows ascending, fixed shuffle tree; pass 2 adds a column's segment sums in order.  Bit-reproducible, no atomics.
// MODE 0: out[v] = sum_i A[i,v] w[i] (slack v: w[v-n]; basic v: 0)     MODE 1: out[v] = |a_v|^2 + 1
constexpr int CSC_SEG = 1024;
// Column-sharded engines price out only the segments [sg0, sg1) of their own column block.
//
// Latency, not bandwidth, bounds this kernel: a column of the config-4 LP holds ~100 entries, so a warp spends its time in
// the dependent chain descriptor -> (row index, value) -> multiplier[row] -> shuffle tree, three global round trips per
// segment (first version: five — segment column, its offset and end, the basic flag, then the entries — 89 us per launch
// = 1.3 TB/s, profiles/r02_price_csc_full.md).  Here a segment is ONE 16-byte descriptor, the basic flag is left to
// k_price_csc_fin, and a warp works on PR_CSC_U segments at once with the first two strides of each in flight together.
// Per segment the sum is unchanged: lane-strided partial sums in ascending entry order, then the fixed shuffle tree.
struct SegDesc {
  int64_t begin;
  int32_t len, col;
};
static_assert(sizeof(SegDesc) == 16, "SegDesc is one 16-byte load");
constexpr int PR_CSC_U = 4;      // short segments (<= 64 entries) in flight per warp
constexpr int PR_CSC_LONG = 64;  // a segment with more entries is "long": one per warp, eight strides in flight
// Where the entries are: the column counts are power-law distributed, so on config 4 ~8 % of the segments (the full
// 1024-entry pieces of the ~1 % longest columns) hold ~85 % of the entries, while ~90 % of the segments are short columns of a
// few dozen entries.  Two work lists (built with the segment table): a LONG segment goes to one warp that keeps eight
// strides (256 entries: values, row indices, then the gathers) in flight — bandwidth; SHORT ones are taken four at a time
// with both strides of each issued together — latency.  Per segment the summation order is the same in both: lane-strided
// partial sums in ascending entry order, then the fixed shuffle tree (bit-identical to the first version of the kernel).
// STREAM: matrix entries are loaded with the evict-first policy (__ldcs: a matrix larger than L2 is streamed once per pivot) or
// with the default policy (the 117 MB of config 4 can stay partly L2-resident between two price-outs; MLP_CSC_STREAM picks).
template <int MODE, bool STREAM>
__global__ void __launch_bounds__(256) k_price_csc_seg(const SegDesc* __restrict__ desc, const int32_t* __restrict__ idx,
                                                       const double* __restrict__ val, const int32_t* __restrict__ long_ids,
                                                       int nlong, const int32_t* __restrict__ short_ids, int nshort,
                                                       const double* __restrict__ w, double* __restrict__ seg_sum) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  for (int i = wid; i < nlong; i += warps) {
    const int sg = long_ids[i];
    const int4 d = __ldg(reinterpret_cast<const int4*>(desc + sg));
    const int64_t b = ((int64_t)(unsigned)d.x) | ((int64_t)d.y << 32);
    const int len = d.z;
    double acc = 0.0;
    for (int o0 = lane; o0 < len; o0 += 256) {
      double a[8];
      int r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int o = o0 + 32 * u;
        const bool ok = o < len;
        a[u] = ok ? (STREAM ? __ldcs(val + b + o) : __ldg(val + b + o)) : 0.0;
        r[u] = (MODE == 0 && ok) ? (STREAM ? __ldcs(idx + b + o) : __ldg(idx + b + o)) : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (o0 + 32 * u < len) acc += (MODE == 0) ? a[u] * w[r[u]] : a[u] * a[u];
    }
    const double tot = warp_sum(acc);
    if (lane == 0) seg_sum[sg] = tot;
  }
  for (int base = wid; base < nshort; base += warps * PR_CSC_U) {
    int64_t b[PR_CSC_U];
    int len[PR_CSC_U], sgq[PR_CSC_U];
#pragma unroll
    for (int q = 0; q < PR_CSC_U; ++q) {
      const int i = base + q * warps;
      b[q] = 0;
      len[q] = 0;
      sgq[q] = -1;
      if (i < nshort) {
        sgq[q] = short_ids[i];
        const int4 d = __ldg(reinterpret_cast<const int4*>(desc + sgq[q]));
        b[q] = ((int64_t)(unsigned)d.x) | ((int64_t)d.y << 32);
        len[q] = d.z;
      }
    }
    double a[PR_CSC_U][2];
    int r[PR_CSC_U][2];
#pragma unroll
    for (int q = 0; q < PR_CSC_U; ++q)
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int o = lane + 32 * it;
        const bool ok = o < len[q];
        a[q][it] = ok ? (STREAM ? __ldcs(val + b[q] + o) : __ldg(val + b[q] + o)) : 0.0;
        r[q][it] = (MODE == 0 && ok) ? (STREAM ? __ldcs(idx + b[q] + o) : __ldg(idx + b[q] + o)) : 0;
      }
#pragma unroll
    for (int q = 0; q < PR_CSC_U; ++q) {
      double acc = 0.0;
#pragma unroll
      for (int it = 0; it < 2; ++it)
        if (lane + 32 * it < len[q]) acc += (MODE == 0) ? a[q][it] * w[r[q][it]] : a[q][it] * a[q][it];
      const double tot = warp_sum(acc);
      if (lane == 0 && sgq[q] >= 0) seg_sum[sgq[q]] = tot;
    }
  }
}
template <int MODE>
__global__ void __launch_bounds__(256) k_price_csc_fin(const int64_t* __restrict__ col_seg, const double* __restrict__ seg_sum,
                                                       int64_t n, int64_t m, int64_t c0, const double* __restrict__ w,
                                                       const uint8_t* __restrict__ vflag, double* __restrict__ out) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // LOCAL variable
  if (v < n) {
    double t = 0.0;
    for (int64_t sg = col_seg[c0 + v]; sg < col_seg[c0 + v + 1]; ++sg) t += seg_sum[sg];
    out[v] = (MODE == 1) ? t + 1.0 : ((vflag[v] & MLP_BASIC) ? 0.0 : t);
  } else if (v < n + m) {
    out[v] = (MODE == 1) ? 2.0 : ((vflag[v] & MLP_BASIC) ? 0.0 : w[v - n]);
  }
}
// rows of A x_N over the CSR copy (solver.rs:234-238): one warp per row
// (a shard sums only the entries of its own column block [c0, c0 + n_loc); xnb is indexed by LOCAL variable)
__global__ void __launch_bounds__(256) k_row_dot_csr(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                     const double* __restrict__ val, int64_t m, int64_t c0, int64_t n_loc,
                                                     const double* __restrict__ xnb, double* __restrict__ out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= m) return;
  double acc = 0.0;
  for (int64_t t = ptr[r] + lane; t < ptr[r + 1]; t += 32) {
    const int64_t j = (int64_t)idx[t] - c0;
    if (j >= 0 && j < n_loc) acc += val[t] * xnb[j];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc;
}

// The three places where the basis machinery touches the basic structural columns, read from the sparse matrix itself
// instead of a dense m x k column cache (12 bytes per stored entry instead of 8 m k):
//   corevar[t]  structural variable of core column t        corepos[v]  core column of variable v, or -1
//   rowcore[i]  core row of constraint row i (i in R), or -1
// FTRAN tail: alpha[cov_i] = a_i - sum_{j in row i, j in the core} A[i,j] x[corepos[j]]  (warp per CSR row)
__global__ void __launch_bounds__(256) k_ftran_finish_csr(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                          const double* __restrict__ val, int m, int k,
                                                          const double* __restrict__ xk, const double* __restrict__ rhs0,
                                                          const int32_t* __restrict__ rowcover, const int32_t* __restrict__ Jpos,
                                                          const int32_t* __restrict__ corepos, double* __restrict__ out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt < k) out[Jpos[gt]] = xk[gt];
  const int64_t i = gt >> 5;
  if (i >= m) return;
  const int cov = rowcover[i];
  if (cov < 0) return;
  double acc = 0.0;
  for (int64_t t = ptr[i] + lane; t < ptr[i + 1]; t += 32) {
    const int c = corepos[idx[t]];
    if (c >= 0) acc += val[t] * xk[c];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[cov] = rhs0[i] - acc;
}
// BTRAN core right-hand side: x[t] = c[Jpos[t]] - sum_i A[i, corevar[t]] cov[i], over the segments of the core columns
// (cseg_id[j] = global segment, cseg_first[t] = first entry of core column t in that list; built at each refactorization)
__global__ void __launch_bounds__(256) k_core_rhs_seg(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                      const double* __restrict__ val, const int32_t* __restrict__ seg_col,
                                                      const int64_t* __restrict__ seg_off, const int32_t* __restrict__ cseg_id,
                                                      int ncseg, const double* __restrict__ cov, double* __restrict__ csum) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= ncseg) return;
  const int64_t sg = cseg_id[j];
  const int v = seg_col[sg];
  const int64_t b = seg_off[sg], e = min(b + (int64_t)CSC_SEG, ptr[v + 1]);
  double acc = 0.0;
  for (int64_t q = b + lane; q < e; q += 32) acc += val[q] * cov[idx[q]];
  acc = warp_sum(acc);
  if (lane == 0) csum[j] = acc;
}
__global__ void k_core_rhs_fin(const double* __restrict__ csum, const int32_t* __restrict__ cseg_first, int k,
                               const double* __restrict__ c, const int32_t* __restrict__ Jpos, double* __restrict__ x) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  double tsum = 0.0;
  for (int j = cseg_first[t]; j < cseg_first[t + 1]; ++j) tsum += csum[j];
  x[t] = c[Jpos[t]] - tsum;
}
// core C = D[R,:]: scatter the stored entries of each core column that fall into core rows (C zero-filled before)
__global__ void __launch_bounds__(256) k_extract_core_seg(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                          const double* __restrict__ val, const int32_t* __restrict__ seg_col,
                                                          const int64_t* __restrict__ seg_off, const int32_t* __restrict__ cseg_id,
                                                          int ncseg, const int32_t* __restrict__ corepos,
                                                          const int32_t* __restrict__ rowcore, double* __restrict__ C, int64_t ld) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= ncseg) return;
  const int64_t sg = cseg_id[j];
  const int v = seg_col[sg];
  const int t = corepos[v];
  const int64_t b = seg_off[sg], e = min(b + (int64_t)CSC_SEG, ptr[v + 1]);
  for (int64_t q = b + lane; q < e; q += 32) {
    const int r = rowcore[idx[q]];
    if (r >= 0) C[(int64_t)t * ld + r] = val[q];
  }
}
__global__ void k_set_corepos(int32_t* __restrict__ corepos, const int32_t* __restrict__ corevar, int k, int clear) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < k) corepos[corevar[t]] = clear ? -1 : t;
}

// ------------------------------------------------------------------------------------------------ K1 pricing scan
// choose_pivot, solver.rs:696-739: arg-max of d^2/gamma (or |d|) over eligible non-basic variables, strict '>' in
// ascending position order => lowest position wins ties.  Writes this shard's candidate header.
__global__ void __launch_bounds__(256) k_select_primal(const double* __restrict__ d, const double* __restrict__ gam,
                                                        const uint8_t* __restrict__ vflag, const int32_t* __restrict__ vpos,
                                                        int64_t nt, int64_t n, int64_t c0, int64_t ng, int use_se,
                                                        double* __restrict__ red_f, long long* __restrict__ red_i,
                                                        unsigned* counter, const double* __restrict__ xnb,
                                                        const int* __restrict__ flags, Cand* out) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  KeyIdx best{-INFINITY, LLONG_MAX};
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double dv = d[v];
    if (((f & MLP_AT_MIN) && dv > -EPS) || ((f & MLP_AT_MAX) && dv < EPS)) continue;  // 705-708
    const double score = use_se ? dv * dv / gam[v] : fabs(dv);
    const long long key2 = ((long long)vpos[v] << 32) | (long long)v;  // position decides ties
    if (better_max(score, key2, best.key, best.idx)) { best.key = score; best.idx = key2; }
  }
  best = block_argmax(best, smk, smi);
  if (threadIdx.x == 0) { red_f[blockIdx.x] = best.key; red_i[blockIdx.x] = best.idx; }
  if (!last_block(counter)) return;
  KeyIdx b{-INFINITY, LLONG_MAX};
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const double k = __ldcg(red_f + q);
    const long long i = __ldcg(red_i + q);
    if (better_max(k, i, b.key, b.idx)) { b.key = k; b.idx = i; }
  }
  b = block_argmax(b, smk, smi);
  if (threadIdx.x == 0) {
    *counter = 0;
    out->f[4] = flags[2] ? 2.0 : (double)flags[0];
    if (b.idx == LLONG_MAX) { out->var = -1; out->key = -INFINITY; out->tie = LLONG_MAX; }
    else {
      const long long v = b.idx & 0xffffffffLL;
      out->key = b.key;
      out->tie = (b.idx >> 32) << 32;
      out->var = v < n ? c0 + v : ng + (v - n);
      out->f[0] = d[v];
      out->f[1] = xnb[v];
    }
  }
}

// low word of the winner's `tie` after absorbing a losing candidate of another shard: that candidate and its own ties
// count towards the winner's when the keys agree exactly (low 16 bits) / within NEAR_TIE (next 16 bits)
__host__ __device__ __forceinline__ long long merge_ties(long long wt, double wk, long long ct, double ck) {
  long long e = wt & 0xffff, n = (wt >> 16) & 0xffff;
  if (ck == wk) e += 1 + (ct & 0xffff);
  if (ck >= wk * (1.0 - 1e-9)) n += 1 + ((ct >> 16) & 0xffff);
  if (e > 65535) e = 65535;
  if (n > 65535) n = 65535;
  return (wt & ~0xffffffffLL) | (n << 16) | e;
}
// Arg-reduce of the gathered candidate headers ON THE DEVICE (larger key wins, ties go to the smaller `tie`: lowest
// position, solver.rs:719) and copy of the winner's column into colq, so that the FTRAN of the entering column can be
// queued behind the selection without a host round trip.  Every thread repeats the <= 8-way comparison.
__global__ void __launch_bounds__(256) k_pick_winner(const char* __restrict__ recv, size_t xbytes, int world, int m,
                                                      double* __restrict__ colq, Cand* __restrict__ win) {
  pdl_wait();
  int best = -1;
  double err = 0.0;
  for (int r = 0; r < world; ++r) {
    const Cand* c = reinterpret_cast<const Cand*>(recv + (size_t)r * xbytes);
    if (c->f[4] != 0.0) err = 1.0;
    if (c->var < 0) continue;
    if (best < 0) { best = r; continue; }
    const Cand* b = reinterpret_cast<const Cand*>(recv + (size_t)best * xbytes);
    if (c->key > b->key || (c->key == b->key && c->tie < b->tie)) best = r;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) colq[i] = best < 0 ? 0.0 : reinterpret_cast<const double*>(recv + (size_t)best * xbytes + sizeof(Cand))[i];
  if (i == 0) {
    if (best < 0) { win->var = -1; win->key = -INFINITY; win->tie = LLONG_MAX; }
    else {
      Cand w = *reinterpret_cast<const Cand*>(recv + (size_t)best * xbytes);
      for (int r = 0; r < world; ++r) {  // ties of the winner across shards (dual ratio test)
        const Cand* c = reinterpret_cast<const Cand*>(recv + (size_t)r * xbytes);
        if (r != best && c->var >= 0) w.tie = merge_ties(w.tie, w.key, c->tie, c->key);
      }
      *win = w;
    }
    win->f[4] = err;
  }
}

// The same exchange as ONE kernel over NVLink peer memory (one process per GPU, buffers mapped with CUDA IPC): every
// rank stores its 64-byte candidate header into a mailbox slot of every peer and raises the slot's sequence flag
// (system-scope fence in between); every rank then polls its OWN mailbox, arg-reduces the headers, and pulls the
// winner's column (8 m bytes) straight out of the owner's memory.  Compared with the all-gather it moves one column
// instead of `world` and has no collective launch latency.  Buffers and mailboxes are double-buffered by exchange
// parity: a rank can be at most one exchange ahead of a peer, because finishing exchange s needs every peer's flag s.
struct PeerTable { char* base[8]; };
constexpr int P2P_SLOT = 128;  // 64 B header + flag, padded
__global__ void __launch_bounds__(256) k_exchange_p2p(PeerTable pt, int rank, int world, unsigned long long seq, int parity,
                                                       const Cand* __restrict__ mine, int m, size_t col_bytes, size_t box_off,
                                                       double* __restrict__ colq, Cand* __restrict__ win) {
  pdl_wait();
  __shared__ Cand hdr[8];
  __shared__ int s_best;
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  if (blockIdx.x == 0 && threadIdx.x < world) {  // publish into peer `threadIdx.x`
    char* slot = pt.base[threadIdx.x] + box_off + ((size_t)parity * world + rank) * P2P_SLOT;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mine);
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(slot);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = src[q];
    __threadfence_system();
    dst[8] = seq;
  }
  __syncthreads();
  if (threadIdx.x < world) {  // wait for peer `threadIdx.x`'s header in my own mailbox
    const char* slot = pt.base[rank] + box_off + ((size_t)parity * world + threadIdx.x) * P2P_SLOT;
    const volatile unsigned long long* src = reinterpret_cast<const volatile unsigned long long*>(slot);
    const long long t0 = clock64();
    bool ok = true;
    while (src[8] != seq) {
      if (clock64() - t0 > 120000000000LL) { ok = false; break; }  // ~60 s: a peer died; report instead of hanging forever
      __nanosleep(64);
    }
    __threadfence_system();
    unsigned long long* d = reinterpret_cast<unsigned long long*>(&hdr[threadIdx.x]);
#pragma unroll
    for (int q = 0; q < 8; ++q) d[q] = src[q];
    if (!ok) s_bad = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = -1;
    for (int r = 0; r < world; ++r) {
      if (hdr[r].var < 0) continue;
      if (best < 0 || hdr[r].key > hdr[best].key || (hdr[r].key == hdr[best].key && hdr[r].tie < hdr[best].tie)) best = r;
    }
    s_best = s_bad ? -1 : best;
  }
  __syncthreads();
  const int best = s_best;
  const double* src = best < 0 ? nullptr : reinterpret_cast<const double*>(pt.base[best] + (size_t)parity * col_bytes);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    colq[i] = best < 0 ? 0.0 : __ldcv(src + i);  // peer memory: never served from a stale cache line
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double err = s_bad ? 2.0 : 0.0;
    for (int r = 0; r < world; ++r) if (hdr[r].f[4] != 0.0 && err == 0.0) err = 1.0;
    if (best < 0) { win->var = -1; win->key = -INFINITY; win->tie = LLONG_MAX; }
    else {
      Cand w = hdr[best];
      for (int r = 0; r < world; ++r)
        if (r != best && hdr[r].var >= 0) w.tie = merge_ties(w.tie, w.key, hdr[r].tie, hdr[r].key);
      *win = w;
    }
    win->f[4] = err;
  }
}

// ------------------------------------------------------------------------------------------------ K11 dual column
// choose_entering_col_dual, solver.rs:919-1021
__device__ __forceinline__ bool dual_eligible(double coeff, unsigned f, int leaving_diff_sign) {
  bool entering_diff_sign;
  if (coeff >= EPS) entering_diff_sign = !leaving_diff_sign;
  else if (coeff <= -EPS) entering_diff_sign = leaving_diff_sign;
  else return false;
  return entering_diff_sign ? !(f & MLP_AT_MAX) : !(f & MLP_AT_MIN);
}
__device__ __forceinline__ double clamp_obj(double oc, unsigned f) {
  if ((f & MLP_AT_MIN) && oc < 0.0) oc = 0.0;
  if ((f & MLP_AT_MAX) && oc > 0.0) oc = 0.0;
  return oc;
}
__global__ void __launch_bounds__(256) k_ratio_dual_1(const double* __restrict__ rc, const double* __restrict__ d,
                                                       const uint8_t* __restrict__ vflag, int64_t nt, int lds,
                                                       double* __restrict__ red_f, unsigned* counter, double* __restrict__ scal) {
  pdl_wait();
  __shared__ double sm[32];
  double best = INFINITY;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double coeff = rc[v];
    if (!dual_eligible(coeff, f, lds)) continue;
    const double oc = clamp_obj(d[v], f);
    const double cur = (fabs(oc) + EPS) / fabs(coeff);  // 970
    if (cur < best) best = cur;
  }
  best = block_min(best, sm);
  if (threadIdx.x == 0) red_f[blockIdx.x] = best;
  if (!last_block(counter)) return;
  double b = INFINITY;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) b = fmin(b, __ldcg(red_f + q));
  b = block_min(b, sm);
  if (threadIdx.x == 0) {
    *counter = 0;
    scal[0] = b;
  }
}
__global__ void k_min_small(const double* __restrict__ vals, int cnt, double* __restrict__ out) {
  pdl_wait();
  double b = INFINITY;
  for (int q = 0; q < cnt; ++q) b = fmin(b, vals[q]);
  *out = b;
}
// pass 2: exact ties in |coeff| go to the lowest GLOBAL variable index (the reference: first-touch order) and
// are counted (candidate header `tie`, low word)
__global__ void __launch_bounds__(256) k_ratio_dual_2(const double* __restrict__ rc, const double* __restrict__ d,
                                                       const uint8_t* __restrict__ vflag, const int32_t* __restrict__ vpos,
                                                       const double* __restrict__ xnb, int64_t nt, int64_t n, int64_t c0,
                                                       int64_t ng, int lds, const double* __restrict__ scal,
                                                       double* __restrict__ red_f, long long* __restrict__ red_i,
                                                       unsigned* counter, const int* __restrict__ flags, Cand* out,
                                                       int scan_slacks) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  __shared__ long long smc[32];
  const double max_step = scal[0];
  KeyIdxC best{-INFINITY, LLONG_MAX, 0, 0};
  // The slack variables are replicated on every shard: only ONE shard (rank 0) proposes and counts them, so that every
  // variable is scanned exactly once across the shards and the tie counts add up to the single-shard ones.
  const int64_t vend = scan_slacks ? nt : n;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < vend; v += (int64_t)gridDim.x * blockDim.x) {
    const unsigned f = vflag[v];
    if (f & MLP_BASIC) continue;
    const double coeff = rc[v];
    if (!dual_eligible(coeff, f, lds)) continue;
    const double oc = clamp_obj(d[v], f);
    const double cur = fabs(oc) / fabs(coeff);  // 993
    const long long g = v < n ? c0 + v : ng + (v - n);
    if (cur <= max_step) kic_merge(best, fabs(coeff), g, 1, 1);
  }
  best = block_argmax_c(best, smk, smi, smc);
  if (threadIdx.x == 0) {
    red_f[blockIdx.x] = best.key;
    red_i[blockIdx.x] = best.idx;
    red_i[RED_CNT_OFF + blockIdx.x] = ((long long)best.ex << 32) | (unsigned)best.nr;
  }
  if (!last_block(counter)) return;
  KeyIdxC b{-INFINITY, LLONG_MAX, 0, 0};
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const double k = __ldcg(red_f + q);
    const long long i = __ldcg(red_i + q);
    const long long c = __ldcg(red_i + RED_CNT_OFF + q);
    kic_merge(b, k, i, (int)(c >> 32), (int)(c & 0xffffffffLL));
  }
  b = block_argmax_c(b, smk, smi, smc);
  if (threadIdx.x == 0) {
    *counter = 0;
    out->f[4] = flags[2] ? 2.0 : (double)flags[0];
    if (b.idx == LLONG_MAX) { out->var = -1; out->key = -INFINITY; out->tie = LLONG_MAX; }
    else {
      const long long g = b.idx;
      const long long v = g >= ng ? n + (g - ng) : g - c0;
      out->key = b.key;
      out->tie = (g << 32) | pack_ties(b.ex, b.nr);
      out->var = g;
      out->f[0] = rc[v];
      out->f[1] = d[v];
      out->f[2] = xnb[v];
      out->f[3] = (double)vpos[v];
    }
  }
}

// ------------------------------------------------------------------------------------------------ pivot updates
// Row half of Solver::pivot: basic values (solver.rs:1049-1055), dual steepest-edge norms (update_dual_sq_norms
// 1163-1173) and the new eta column (push_eta_matrix 1274-1284).  Replicated on every shard.
__global__ void __launch_bounds__(256) k_pivot_rows(const double* __restrict__ alpha, const double* __restrict__ tau,
                                                     double* __restrict__ xB, double* __restrict__ w, int m, int row,
                                                     double entering_new_val, double entering_diff, double coeff, int has_elem,
                                                     int dse, const double* __restrict__ scal, double* __restrict__ eta_col,
                                                     int* __restrict__ flags, uint8_t* __restrict__ touched,
                                                     const uint8_t* __restrict__ touched_new) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  if (eta_col) touched[r] = touched_new[r];  // the pushed eta stores every listed position of col_coeffs (solver.rs:1274-1284)
  const double a = alpha[r];
  if (!has_elem) {  // bound flip, solver.rs:1035-1037
    if (a != 0.0) xB[r] -= entering_diff * a;
    return;
  }
  if (r == row) xB[r] = entering_new_val;
  else if (a != 0.0) xB[r] -= entering_diff * a;
  if (dse) {
    const double pivot_sq_norm = scal[1];  // |rho|^2, solver.rs:1160
    const double pcs = coeff * coeff;
    if (r == row) {
      w[r] = pivot_sq_norm / pcs;
      if (!isfinite(w[r])) flags[0] = 1;
    } else if (a != 0.0) {
      const double nw = w[r] + (-2.0 * a * tau[r] / coeff + pivot_sq_norm * a * a / pcs);  // 1168-1169
      w[r] = nw;
      if (!isfinite(nw)) flags[0] = 1;
    }
  }
  if (eta_col) eta_col[r] = (r == row) ? 1.0 - 1.0 / coeff : a / coeff;  // 1276-1280
}
__global__ void k_flip_var(double* xnb, uint8_t* vflag, const double* lo, const double* hi, int64_t q, int64_t ql, double new_val) {
  pdl_wait();
  xnb[ql] = new_val;
  unsigned f = vflag[ql] & MLP_FIXED;
  if (new_val == lo[q]) f |= MLP_AT_MIN;
  if (new_val == hi[q]) f |= MLP_AT_MAX;
  vflag[ql] = (uint8_t)f;  // solver.rs:1038-1040
}

// The variable half of Solver::pivot and the NEXT pricing scan in one pass over this shard's variables (SURVEY K1
// "fused update+select"): per variable — finish the N^T v price-out (sum of the chunk partials in chunk order, as
// k_price_finish), update reduced cost and primal steepest-edge norm (k_pivot_vars), apply the basis swap to the two
// variables concerned (k_pivot_swap), then score the variable for choose_pivot (k_select_primal) with its new state.
// Block partial arg-max -> last block finishes and leaves the candidate header for the exchange step.
struct UpdSel {
  // price-out finish (pse only; sparse storage has helper already)
  const double* partial; const int32_t* count_ptr; int64_t lda; const double* slack_vals; double* helper; int finish;
  // pivot
  int64_t q, ql, lvl, lv; int col, row; double pivot_obj, coeff, leaving_new_val; int pse;
};
__global__ void __launch_bounds__(256) k_update_select(UpdSel a, double* __restrict__ d, double* __restrict__ gam,
                                                        const double* __restrict__ rc, double* __restrict__ xnb,
                                                        uint8_t* __restrict__ vflag, int32_t* __restrict__ vpos,
                                                        int32_t* __restrict__ bvar, double* __restrict__ loB,
                                                        double* __restrict__ hiB, const double* __restrict__ lo,
                                                        const double* __restrict__ hi, int64_t nt, int64_t n, int64_t m,
                                                        int64_t c0, int64_t ng, const double* __restrict__ scal,
                                                        int* __restrict__ flags, double* __restrict__ red_f,
                                                        long long* __restrict__ red_i, unsigned* counter, DevRes* res, Cand* out) {
  pdl_wait();
  __shared__ double smk[32];
  __shared__ long long smi[32];
  KeyIdx best{-INFINITY, LLONG_MAX};
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nt) {
    unsigned f = vflag[v];
    double h = 0.0;
    if (a.pse) {
      if (a.finish) {
        if (v < n) {
          const int C = price_chunks_for(*a.count_ptr);
          for (int c = 0; c < C; ++c) h += a.partial[(int64_t)c * a.lda + v];
        } else h = a.slack_vals[v - n];
        if (f & MLP_BASIC) h = 0.0;
        a.helper[v] = h;
      } else h = a.helper[v];
    }
    double dv = d[v], gv = gam[v];
    if (v == a.lvl) {  // the leaving variable takes the non-basic slot (solver.rs:1066-1071, 1076, 1142)
      xnb[v] = a.leaving_new_val;
      f = 0;
      if (a.leaving_new_val == lo[a.lv]) f |= MLP_AT_MIN;
      if (a.leaving_new_val == hi[a.lv]) f |= MLP_AT_MAX;
      vflag[v] = (uint8_t)f;
      vpos[v] = a.col;
      dv = -a.pivot_obj;
      d[v] = dv;
      if (a.pse) {
        gv = (scal[2] + 1.0) / (a.coeff * a.coeff);
        gam[v] = gv;
        if (!isfinite(gv)) flags[0] = 1;
      }
    } else if (v == a.ql) {  // the entering variable becomes basic (1088-1091)
      f = MLP_BASIC;
      vflag[v] = MLP_BASIC;
      vpos[v] = a.row;
    } else if (!(f & MLP_BASIC)) {
      const double c = rc[v];
      if (c != 0.0) {
        dv -= a.pivot_obj * c;  // 1073-1080
        d[v] = dv;
        if (a.pse) {
          const double psn = scal[2] + 1.0;  // 1136
          gv = gv + (-2.0 * c * h / a.coeff + psn * c * c / (a.coeff * a.coeff));  // 1144-1146
          gam[v] = gv;
          if (!isfinite(gv)) flags[0] = 1;
        }
      }
    }
    if (v == 0) {  // row-side bookkeeping, identical on every shard (1057-1058, 1088)
      loB[a.row] = lo[a.q];
      hiB[a.row] = hi[a.q];
      res->i[0] = bvar[a.row];  // the device's idea of the leaving variable, cross-checked by the host
      bvar[a.row] = (int32_t)a.q;
    }
    // choose_pivot's scan (696-739) on the updated state
    if (!(f & MLP_BASIC) && !(((f & MLP_AT_MIN) && dv > -EPS) || ((f & MLP_AT_MAX) && dv < EPS))) {
      best.key = a.pse ? dv * dv / gv : fabs(dv);
      best.idx = ((long long)vpos[v] << 32) | (long long)v;
    }
  }
  best = block_argmax(best, smk, smi);
  if (threadIdx.x == 0) { red_f[blockIdx.x] = best.key; red_i[blockIdx.x] = best.idx; }
  if (!last_block(counter)) return;
  KeyIdx b{-INFINITY, LLONG_MAX};
  for (int qd = threadIdx.x; qd < (int)gridDim.x; qd += blockDim.x) {
    const double k = __ldcg(red_f + qd);
    const long long i = __ldcg(red_i + qd);
    if (better_max(k, i, b.key, b.idx)) { b.key = k; b.idx = i; }
  }
  b = block_argmax(b, smk, smi);
  if (threadIdx.x == 0) {
    *counter = 0;
    const int nf = *((volatile int*)flags);
    res->flags[0] = nf;
    res->flags[1] = flags[1];
    out->f[4] = flags[2] ? 2.0 : (double)nf;
    if (b.idx == LLONG_MAX) { out->var = -1; out->key = -INFINITY; out->tie = LLONG_MAX; }
    else {
      const long long vv = b.idx & 0xffffffffLL;
      out->key = b.key;
      out->tie = (b.idx >> 32) << 32;
      out->var = vv < n ? c0 + vv : ng + (vv - n);
      out->f[0] = __ldcg(d + vv);
      out->f[1] = __ldcg(xnb + vv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ init kernels
// partial of A x_N over this shard's columns (solver.rs:234-238). One CTA per row.
__global__ void __launch_bounds__(256) k_row_dot(const double* __restrict__ A, int64_t lda, int64_t n,
                                                  const double* __restrict__ xnb, double* __restrict__ out) {
  pdl_wait();
  __shared__ double sm[32];
  const int r = blockIdx.x;
  const double* row = A + (int64_t)r * lda;
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += row[j] * xnb[j];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) out[r] = tot;
}
// basic_var_vals = rhs - sum over shards (in rank order) of the partial products
__global__ void k_init_basic_vals(const double* __restrict__ parts, int world, int m, const double* __restrict__ rhs,
                                  double* __restrict__ xB) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double tot = 0.0;
  for (int g = 0; g < world; ++g) tot += parts[(int64_t)g * m + r];
  xB[r] = rhs[r] - tot;
}
// d_N = c_N - N^T y (recalc_obj_coeffs, solver.rs:1216-1222)
__global__ void k_recalc_d(const double* __restrict__ cobj, const double* __restrict__ rc, const uint8_t* __restrict__ vflag,
                           int64_t nt, int64_t n, int64_t c0, int64_t ng, double* __restrict__ d) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nt || (vflag[v] & MLP_BASIC)) return;
  const int64_t g = v < n ? c0 + v : ng + (v - n);
  d[v] = cobj[g] - rc[v];
}
// objective from scratch (solver.rs:1224-1230) in three parts: basic rows, non-basic slacks (both replicated),
// non-basic structurals of this shard.  Single CTA, deterministic.
__global__ void __launch_bounds__(1024) k_recalc_obj(const double* __restrict__ cobj, const int32_t* __restrict__ bvar,
                                                      const double* __restrict__ xB, int m, const double* __restrict__ xnb,
                                                      const uint8_t* __restrict__ vflag, int64_t n, int64_t c0, int64_t ng,
                                                      double* __restrict__ out3) {
  pdl_wait();
  __shared__ double sm[32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int r = threadIdx.x; r < m; r += blockDim.x) a += cobj[bvar[r]] * xB[r];
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x)
    if (!(vflag[n + i] & MLP_BASIC)) b += cobj[ng + i] * xnb[n + i];
  for (int64_t v = threadIdx.x; v < n; v += blockDim.x)
    if (!(vflag[v] & MLP_BASIC)) c += cobj[c0 + v] * xnb[v];
  const double ta = block_sum(a, sm);
  const double tb = block_sum(b, sm);
  const double tc = block_sum(c, sm);
  if (threadIdx.x == 0) { out3[0] = ta; out3[1] = tb; out3[2] = tc; }
}
__global__ void k_gather_cB(const double* __restrict__ cobj, const int32_t* __restrict__ bvar, int m, double* __restrict__ out) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m) out[r] = cobj[bvar[r]];
}

// ------------------------------------------------------------------------------------------------ incremental API (row f2)
// Solver::add_constraint (solver.rs:549-634) pieces.  A cut may carry coefficients g_i on slack variables
// (add_gomory_cut, 440-460); slack columns stay unit columns here, so s_i = rhs_i - a_i x is substituted:
// row' = c - A^T g, rhs' = rhs - g . rhs_old (the same constraint; see DESIGN.md §8 for what that changes).
__global__ void k_row_combine(double* __restrict__ row, const double* __restrict__ partial, const int32_t* __restrict__ count_ptr,
                              int64_t lda, int64_t n) {
  pdl_wait();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int C = price_chunks_for(*count_ptr);
  double t = 0.0;
  for (int c = 0; c < C; ++c) t += partial[(int64_t)c * lda + j];
  row[j] -= t;
}
// out[0] = base - sum_i a_i b_i (single CTA, deterministic); used for rhs' and for the new basic value rhs - a . x
__global__ void __launch_bounds__(1024) k_sub_dot(const double* __restrict__ a, const double* __restrict__ b, int64_t cnt, double base,
                                                   const double* __restrict__ base_ptr, double* __restrict__ out) {
  pdl_wait();
  __shared__ double sm[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) acc += a[i] * b[i];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) out[0] = (base_ptr ? *base_ptr : base) - tot;
}
// current value of every structural variable (Solver::get_value, 371-376) as a dense vector
__global__ void k_struct_values(const double* __restrict__ xnb, const double* __restrict__ xB, const uint8_t* __restrict__ vflag,
                                const int32_t* __restrict__ vpos, int64_t n, double* __restrict__ out) {
  pdl_wait();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = (vflag[j] & MLP_BASIC) ? xB[vpos[j]] : xnb[j];
}
// state of the appended row and of its slack variable (563-571, 591)
__global__ void k_new_row_state(int64_t r, int64_t lv, int64_t gv, double smin, double smax, const double* __restrict__ val,
                                const double* __restrict__ rhs_new, double* lo, double* hi, double* cobj, double* d, double* gam,
                                double* xnb, uint8_t* vflag, int32_t* vpos, int32_t* bvar, double* xB, double* loB, double* hiB,
                                double* w, double* rhs, int32_t* rowcover) {
  pdl_wait();
  lo[gv] = smin; hi[gv] = smax; cobj[gv] = 0.0;
  d[lv] = 0.0; gam[lv] = 0.0; xnb[lv] = 0.0;
  vflag[lv] = MLP_BASIC; vpos[lv] = (int32_t)r;
  bvar[r] = (int32_t)gv; xB[r] = *val; loB[r] = smin; hiB[r] = smax; w[r] = 1.0; rhs[r] = *rhs_new;
  rowcover[r] = (int32_t)r;
}
// the cached basis columns get their entry of the new row
__global__ void k_bcols_new_row(const double* __restrict__ rowA, const int32_t* __restrict__ slots, const int32_t* __restrict__ vars,
                                int cnt, int64_t ldb, int64_t r, double* __restrict__ Bcols) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) Bcols[(int64_t)slots[t] * ldb + r] = rowA[vars[t]];
}
// primal_edge_sq_norms[c] += coeff^2 over the new tableau row (618-622)
__global__ void k_add_sq(double* __restrict__ gam, const double* __restrict__ rc, const uint8_t* __restrict__ vflag, int64_t nt) {
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nt && !(vflag[v] & MLP_BASIC)) gam[v] += rc[v] * rc[v];
}
__global__ void k_copy1(double* dst, const double* src) {
  pdl_wait(); *dst = *src; }
__global__ void k_set_var_state(uint8_t* vflag, int64_t lv, unsigned f) {
  pdl_wait(); vflag[lv] = (uint8_t)f; }

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

// Column cache / LU arenas.  First allocation is generous (~1 GB of basis columns): cudaFree/cudaMalloc of the big
// arenas costs tens of milliseconds, so capacity grows by doubling and rarely; the cache content survives growth.
static mlp_status ensure_lu_capacity(mlp_engine* e, int64_t k, bool exact = false) {
  if (k <= e->kcap && e->Bcols) return MLP_OK;
  int64_t cap = std::max<int64_t>(e->kcap, std::min<int64_t>(e->m, std::max<int64_t>(64, std::min<int64_t>(1024, (1ll << 30) / (8 * e->mld)))));
  // sparse storage keeps no column cache: the arenas are kcap^2 (factors, inverse) — start at 4096 columns (0.27 GB) so that the
  // first thousands of pivots meet no growth (each growth is a re-allocation AND a true factorization: 5 - 20 ms on config 4)
  if (e->sparse) cap = std::max<int64_t>(e->kcap, std::min<int64_t>(e->m, 4096));
  while (cap < k) cap *= 2;  // may exceed m: slots of columns that left since the last refactor stay occupied
  if (exact) cap = k;        // clone: same leading dimensions as the source
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  PoolScope pool(e->use_pool ? e->stream : nullptr);
  double* nb = nullptr;
  ST(dev_alloc(&nb, e->sparse ? 1 : (size_t)e->mld * cap));  // sparse storage reads the basic columns from the matrix itself
  if (!e->sparse && e->Bcols && e->kcap > 0) {
    CU(cudaMemcpyAsync(nb, e->Bcols, (size_t)e->mld * e->kcap * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  for (int64_t s = cap - 1; s >= e->kcap; --s) e->h_free_slots.push_back((int32_t)s);
  dev_free(e->Jpos); dev_free(e->Jslot); dev_free(e->Rp); dev_free(e->Bcols); dev_free(e->LUc); dev_free(e->Cinv);
  dev_free(e->lu_aff); dev_free(e->lu_perm); dev_free(e->lu_rcnt);
  ST(dev_alloc(&e->lu_aff, 192)); ST(dev_alloc(&e->lu_perm, cap)); ST(dev_alloc(&e->lu_rcnt, cap));
  if (e->sparse) {
    if (e->corevar_k > 0) {  // un-mark with the old list before it is freed
      LAUNCH(e, k_set_corepos, cdiv(e->corevar_k, 256), 256, 0, e->corepos, e->corevar, (int)e->corevar_k, 1);
      CU(cudaStreamSynchronize(e->stream));
      e->corevar_k = 0;
    }
    dev_free(e->corevar); dev_free(e->cseg_first);
    ST(dev_alloc(&e->corevar, cap)); ST(dev_alloc(&e->cseg_first, cap + 1));
    dev_free(e->rf_map); dev_free(e->rf_W); dev_free(e->rf_T); dev_free(e->rf_Ep);
    ST(dev_alloc(&e->rf_map, 3 * (size_t)cap + 2 * RF_CAP));
    ST(dev_alloc(&e->rf_W, (size_t)RF_CAP * cap)); ST(dev_alloc(&e->rf_T, (size_t)RF_CAP * cap)); ST(dev_alloc(&e->rf_Ep, (size_t)RF_CAP * cap));
    e->inv_valid = false;  // C^-1 does not survive the re-allocation: the next refactorization is a true one
  }
  e->Bcols = nb;
  e->kcap = cap;
  ST(dev_alloc(&e->Jpos, cap)); ST(dev_alloc(&e->Jslot, cap)); ST(dev_alloc(&e->Rp, cap));
  ST(dev_alloc(&e->LUc, (size_t)cap * cap)); ST(dev_alloc(&e->Cinv, (size_t)cap * cap));
  for (int l = 0; l < 2; ++l) {
    Lane& ln = e->lane[l];
    dev_free(ln.xk); dev_free(ln.xk2); dev_free(ln.gt_part_k);
    ST(dev_alloc(&ln.xk, cap)); ST(dev_alloc(&ln.xk2, cap)); ST(dev_alloc(&ln.gt_part_k, (size_t)GT_MAXSPLIT * cap));
  }
  return MLP_OK;
}
static mlp_status ensure_eta_capacity(mlp_engine* e, int64_t K, bool exact = false) {
  if (K <= e->Kcap && e->E) return MLP_OK;
  int64_t cap = std::max<int64_t>(e->Kcap, std::max<int64_t>(96, std::min<int64_t>(2080, (2ll << 30) / (8 * e->mld))));
  while (cap < K) cap *= 2;
  if (exact) cap = K;
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  PoolScope pool(e->use_pool ? e->stream : nullptr);
  dev_free(e->E); dev_free(e->Ginv); dev_free(e->gK); dev_free(e->etaR); dev_free(e->etaPrev); dev_free(e->etaHead);
  e->Kcap = cap;
  ST(dev_alloc(&e->E, (size_t)e->mld * cap)); ST(dev_alloc(&e->Ginv, (size_t)cap * cap)); ST(dev_alloc(&e->gK, cap));
  ST(dev_alloc(&e->etaR, cap)); ST(dev_alloc(&e->etaPrev, cap)); ST(dev_alloc(&e->etaHead, cap));
  for (int l = 0; l < 2; ++l) {
    Lane& ln = e->lane[l];
    dev_free(ln.tK); dev_free(ln.tK2); dev_free(ln.gt_part_K);
    ST(dev_alloc(&ln.tK, cap)); ST(dev_alloc(&ln.tK2, cap)); ST(dev_alloc(&ln.gt_part_K, (size_t)GT_MAXSPLIT * cap));
  }
  return MLP_OK;
}

// Compact row-major copy of the k basic structural columns (core column ids), built on the device from the CSC copy with the
// same segmented counting transpose as the CSC copy itself (sparse_build.cuh; input "rows" = the core columns in core
// order, so every row of the copy lists its entries in ascending core column).  jvar: the core columns' variables, already
// uploaded to e->corevar.
constexpr int DCSR_CHUNKS = 128;
static mlp_status build_core_rows(mlp_engine* e, const std::vector<int32_t>& jvar) {
  const int64_t m = e->m, k = (int64_t)jvar.size();
  PoolScope pool(e->use_pool ? e->stream : nullptr);
  if (!e->dcsr_ptr) {
    ST(dev_alloc(&e->dcsr_ptr, (size_t)e->mld + 1));
    ST(dev_alloc(&e->dcsr_hist, (size_t)DCSR_CHUNKS * e->mld));
    ST(dev_alloc(&e->dcsr_cnt, (size_t)e->mld));
  }
  if (k == 0) {
    CU(cudaMemsetAsync(e->dcsr_ptr, 0, (size_t)(m + 1) * sizeof(int64_t), e->stream));
    return MLP_OK;
  }
  int64_t nz = 0;
  for (int32_t v : jvar) nz += e->h_csc_ptr[(size_t)v + 1] - e->h_csc_ptr[(size_t)v];
  if (nz > e->dcsr_cap) {
    CU(cudaStreamSynchronize(e->lane[1].st));
    dev_free(e->dcsr_idx); dev_free(e->dcsr_val);
    e->dcsr_cap = std::max<int64_t>(4 * nz, 1 << 22);  // 12 bytes per entry: grow rarely
    ST(dev_alloc(&e->dcsr_idx, (size_t)e->dcsr_cap)); ST(dev_alloc(&e->dcsr_val, (size_t)e->dcsr_cap));
  }
  const int ncseg = (int)e->ncseg;                                   // the core's segments (e->cseg_id), in core-column order
  const int spc = (ncseg + DCSR_CHUNKS - 1) / DCSR_CHUNKS;          // segments per chunk
  const int chunks = (ncseg + spc - 1) / spc;
  CU(cudaMemsetAsync(e->dcsr_hist, 0, (size_t)chunks * m * sizeof(int32_t), e->stream));
  LAUNCH(e, k_d_hist, cdiv((int64_t)ncseg * 32, 256), 256, 0, (const int4*)e->seg_desc, e->cseg_id, ncseg, e->csc_idx, m, spc, e->dcsr_hist);
  // per constraint row: scan over the chunks + row counts (k_t_colscan's segment output is not needed: lane scratch)
  LAUNCH(e, k_t_colscan, cdiv(m, 256), 256, 0, e->dcsr_hist, m, chunks, e->dcsr_cnt, (int64_t*)e->lane[0].wm, CSC_SEG);
  LAUNCH(e, k_scan_excl, 1, 1024, 0, e->dcsr_cnt, m, e->dcsr_ptr);
  LAUNCH(e, k_d_fill, chunks, 256, 0, (const int4*)e->seg_desc, e->cseg_id, ncseg, e->csc_idx, e->csc_val, m, spc, e->corepos, e->dcsr_hist,
         e->dcsr_ptr, e->dcsr_idx, e->dcsr_val);
  return MLP_OK;
}

static void refac_stage(mlp_engine* e, const char* name) {
  if (!e->refac_trace) return;
  if (e->refac_trace == 1) for (int l = 0; l < 2; ++l) cudaStreamSynchronize(e->lane[l].st);  // 2: host-side times only, no extra syncs
  const auto now = std::chrono::steady_clock::now();
  if (name) {
    const double ms = std::chrono::duration<double, std::milli>(now - e->refac_t).count();
    if (e->refac_trace == 2 && ms > 0.5)
      fprintf(stderr, "[refactor event] #%lld since-lu %lld k %lld K %lld: %.3f ms in '%s'\n", (long long)e->cnt.refactors,
              (long long)e->pivots_since_lu, (long long)e->k, (long long)e->K, ms, name);
    bool found = false;
    for (auto& st : e->refac_stage) if (st.first == name) { st.second += ms; found = true; break; }
    if (!found) e->refac_stage.emplace_back(name, ms);
  }
  e->refac_t = std::chrono::steady_clock::now();
}
static void refac_report(mlp_engine* e) {
  if (!e->refac_trace || e->cnt.refactors == 0) return;
  double tot = 0.0;
  for (auto& st : e->refac_stage) tot += st.second;
  fprintf(stderr, "[refactor trace] %lld refactorizations (%lld of them product-form refreshes), mean k %.0f, %.3f ms each\n",
          (long long)e->cnt.refactors, (long long)e->cnt.refreshes, e->refac_k_sum / e->cnt.refactors, tot / e->cnt.refactors);
  fprintf(stderr, "[refactor trace] host blocked in %lld per-pivot device waits: %.1f ms in total (%.1f us each) over %lld basis changes, %lld launches\n",
          (long long)e->waits, e->wait_ms, e->waits ? 1e3 * e->wait_ms / e->waits : 0.0, (long long)e->pivot_seq, (long long)e->cnt.kernel_launches);
  fprintf(stderr, "[refactor trace] accuracy probe (normwise backward error of sampled columns of C^-1): worst accepted refresh %.3g, "
                  "worst after a true factorization %.3g, tolerance %.3g, rejected refreshes %lld\n", e->rf_worst, e->rf_worst_true, e->rf_tol,
          (long long)e->cnt.refresh_rejects);
  for (auto& st : e->refac_stage) fprintf(stderr, "[refactor trace]   %-28s %9.3f ms each  %5.1f %%\n", st.first, st.second / e->cnt.refactors, 100.0 * st.second / tot);
}

// Product-form refresh (refresh_inverse.cuh): C_new^-1 from C_old^-1 and the eta file, written into the LUc buffer, which
// then becomes Cinv.  jpos / R: the NEW core's positions and rows.  Runs before anything of the old factor state (index maps,
// compact core rows, eta file) is touched; both lanes are drained.
// Pinned staging: every index array of a refactorization goes through ONE pinned buffer (a cudaMemcpyAsync from pageable memory
// first waits for the stream and then copies synchronously: eight of them serialised the host with the device).
static mlp_status stage_begin(mlp_engine* e, size_t ints_needed) {
  e->stg_cur ^= 1;
  const int c = e->stg_cur;
  if (!e->stg_ev[c]) CU(cudaEventCreateWithFlags(&e->stg_ev[c], cudaEventDisableTiming));
  else CU(cudaEventSynchronize(e->stg_ev[c]));  // the copies that last used this buffer (two refactorizations ago) are long done
  if (ints_needed > e->stg_cap[c]) {
    if (e->stg_h[c]) cudaFreeHost(e->stg_h[c]);
    e->stg_h[c] = nullptr;
    e->stg_cap[c] = std::max<size_t>(2 * ints_needed, (size_t)1 << 16);
    CU(cudaHostAlloc((void**)&e->stg_h[c], e->stg_cap[c] * sizeof(int32_t), cudaHostAllocDefault));
  }
  e->stg_off = 0;
  return MLP_OK;
}
static mlp_status stage_put(mlp_engine* e, void* dst_dev, const int32_t* src, size_t n) {
  if (n == 0) return MLP_OK;
  const int c = e->stg_cur;
  if (e->stg_off + n > e->stg_cap[c]) { set_err("refactor: staging buffer too small"); return MLP_INVALID; }
  int32_t* h = e->stg_h[c] + e->stg_off;
  std::memcpy(h, src, n * sizeof(int32_t));
  e->stg_off += n;
  CU(cudaMemcpyAsync(dst_dev, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  e->cnt.h2d_bytes += (int64_t)(n * sizeof(int32_t));
  return MLP_OK;
}
static mlp_status stage_end(mlp_engine* e) {
  CU(cudaEventRecord(e->stg_ev[e->stg_cur], e->stream));
  return MLP_OK;
}

// dst[idx[i]] = val[i]
__global__ void k_patch_i32(int32_t* __restrict__ dst, const int32_t* __restrict__ idx, const int32_t* __restrict__ val, int n) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = val[i];
}
// rowcore[Rp[c]] = c
__global__ void k_set_rowcore(int32_t* __restrict__ rowcore, const int32_t* __restrict__ Rp, int k) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < k) rowcore[Rp[c]] = c;
}

// The index sets of the new basis derived from those of the factorized one and the list of basis changes since
// (sparse storage, every change recorded): O(k + K log k) instead of three passes over all m positions / rows.
//   jpos   core columns' positions in order_simple's order (ordering.rs:4-21: ascending entry count, ascending position within a count)
//   R      core rows, ascending
//   prow / pval   rows whose rowcover entry changes and the new values (device patch)
//   gone   rows that left the core (rowcore <- -1)
// e->h_rowcover_f is updated in place.
static mlp_status incremental_sets(mlp_engine* e, std::vector<int32_t>& jpos, std::vector<int32_t>& R, std::vector<int32_t>& prow,
                                   std::vector<int32_t>& pval, std::vector<int32_t>& gone) {
  const int64_t ng = e->ng;
  std::vector<int32_t> P;       // distinct changed positions
  std::vector<int64_t> oldv;    // variable the factorized basis held there
  for (size_t j = 0; j < e->h_eta_pos.size(); ++j) {
    const int32_t p = e->h_eta_pos[j];
    if (std::find(P.begin(), P.end(), p) == P.end()) { P.push_back(p); oldv.push_back(e->h_eta_leave[j]); }
  }
  std::vector<int32_t>& rc = e->h_rowcover_f;
  std::vector<int32_t> T;  // slack rows involved
  e->h_rc_old_rows.clear();
  e->h_rc_old_vals.clear();
  auto touch = [&](int32_t i) {
    if (std::find(T.begin(), T.end(), i) != T.end()) return;
    T.push_back(i);
    e->h_rc_old_rows.push_back(i);
    e->h_rc_old_vals.push_back(rc[(size_t)i]);  // the factorized basis' value: the refresh needs it
  };
  for (size_t q = 0; q < P.size(); ++q) if (oldv[q] >= ng) touch((int32_t)(oldv[q] - ng));
  for (size_t q = 0; q < P.size(); ++q) { const int64_t nv = e->h_bvar[(size_t)P[q]]; if (nv >= ng) touch((int32_t)(nv - ng)); }
  for (size_t q = 0; q < P.size(); ++q) if (oldv[q] >= ng) rc[(size_t)(oldv[q] - ng)] = -1;
  for (size_t q = 0; q < P.size(); ++q) { const int64_t nv = e->h_bvar[(size_t)P[q]]; if (nv >= ng) rc[(size_t)(nv - ng)] = P[q]; }
  auto key_less = [&](int32_t pa, int32_t pb) {  // order_simple's key of the column at a position
    const int64_t va = e->h_bvar[(size_t)pa], vb = e->h_bvar[(size_t)pb];
    const int64_t ca = e->h_csc_ptr[(size_t)va + 1] - e->h_csc_ptr[(size_t)va], cb = e->h_csc_ptr[(size_t)vb + 1] - e->h_csc_ptr[(size_t)vb];
    return ca != cb ? ca < cb : pa < pb;
  };
  jpos.clear();
  jpos.reserve(e->h_Jpos_f.size() + P.size());
  std::vector<int32_t> Ps(P);
  std::sort(Ps.begin(), Ps.end());
  for (int32_t p : e->h_Jpos_f)
    if (!std::binary_search(Ps.begin(), Ps.end(), p)) jpos.push_back(p);  // unchanged columns keep their relative order
  for (size_t q = 0; q < P.size(); ++q) {
    if (e->h_bvar[(size_t)P[q]] >= ng) continue;
    jpos.insert(std::lower_bound(jpos.begin(), jpos.end(), P[q], key_less), P[q]);
  }
  R = e->h_R_sorted;
  prow.clear(); pval.clear(); gone.clear();
  for (int32_t i : T) {
    prow.push_back(i);
    pval.push_back(rc[(size_t)i]);
    auto it = std::lower_bound(R.begin(), R.end(), i);
    const bool in_old = it != R.end() && *it == i;
    const bool in_new = rc[(size_t)i] < 0;
    if (in_old && !in_new) { R.erase(it); gone.push_back(i); }
    else if (!in_old && in_new) R.insert(it, i);
  }
  return MLP_OK;
}

// k_rf_probe on the current C^-1 against the compact rows of the current basic columns; one read-back.  *err: the largest
// normwise backward error over the sampled columns, *core_entries: entries of the core.
static mlp_status probe_inverse(mlp_engine* e, int64_t k, double* err, int64_t* core_entries) {
  unsigned long long* out = e->d_nnzcnt;
  CU(cudaMemsetAsync(out, 0, (1 + 2 * RF_PROBE) * sizeof(unsigned long long), e->stream));
  const int ncol = (int)std::min<int64_t>(RF_PROBE, k);
  LAUNCH(e, k_rf_probe, cdiv(k, 256), 256, 0, e->dcsr_ptr, e->dcsr_idx, e->dcsr_val, e->Rp, (int)k, e->Cinv, e->kcap,
         (int)((e->cnt.refactors * 2654435761ull) % (unsigned long long)k), (int)std::max<int64_t>(1, k / RF_PROBE), ncol, out);
  unsigned long long h[1 + 2 * RF_PROBE];
  ST(d2h(e, h, out, sizeof(h)));
  *core_entries = (int64_t)h[0];
  double worst = 0.0;
  for (int q = 0; q < ncol; ++q) {
    double num, den;
    std::memcpy(&num, &h[1 + q], 8);
    std::memcpy(&den, &h[1 + RF_PROBE + q], 8);
    const double r = den > 0.0 ? num / den : num;
    if (!(r <= worst)) worst = r;
  }
  *err = worst;
  return MLP_OK;
}
static bool can_refresh(const mlp_engine* e) {
  // the rank-K product costs 2 k^2 K flops: worth it while the eta file is short next to the core (a factorization is ~2 k^3)
  return e->sparse && e->inv_valid && e->lu_every > 0 && e->k > 0 && e->K >= 1 && e->K <= std::min<int64_t>(e->Kcap, RF_CAP) &&
         e->K <= std::max<int64_t>(RF_MAXK, e->k / 2) &&
         (int64_t)e->h_Jpos_f.size() == e->k && (int64_t)e->h_eta_pos.size() == e->K && (int64_t)e->h_pos_core.size() == e->m &&
         e->rf_map != nullptr;
}
static mlp_status refresh_inverse(mlp_engine* e, const std::vector<int32_t>& jpos, const std::vector<int32_t>& R) {
  const int k_old = (int)e->k, K = (int)e->K, k_new = (int)jpos.size();
  const int64_t ld = e->kcap;
  std::vector<int32_t> qpos, wrow;  // old slack positions whose row of B_old^-1 is needed, and the rows of those slacks
  auto q_of = [&](int32_t p) -> int {
    for (size_t q = 0; q < qpos.size(); ++q) if (qpos[q] == p) return (int)q;
    return -1;
  };
  std::vector<int32_t> map((size_t)3 * k_new + 2 * (size_t)K + 8, 0);
  int32_t *rowsrc = map.data(), *colsrc = rowsrc + k_new, *jposn = colsrc + k_new, *etasrc = jposn + k_new, *wr = etasrc + K;
  for (int j = 0; j < K; ++j) {
    const int32_t p = e->h_eta_pos[(size_t)j];
    if (e->h_pos_core[(size_t)p] >= 0) { etasrc[j] = e->h_pos_core[(size_t)p]; continue; }
    int q = q_of(p);
    if (q < 0) {  // the FIRST eta at a position tells which variable the factorized basis held there
      const int64_t v = e->h_eta_leave[(size_t)j];
      if (v < e->ng) { set_err("refresh: basis bookkeeping inconsistent (structural variable at a slack position)"); return MLP_INVALID; }
      q = (int)qpos.size();
      qpos.push_back(p);
      wrow.push_back((int32_t)(v - e->ng));
    }
    etasrc[j] = -1 - q;
  }
  for (int t = 0; t < k_new; ++t) {
    const int32_t p = jpos[(size_t)t];
    jposn[t] = p;
    if (e->h_pos_core[(size_t)p] >= 0) { rowsrc[t] = e->h_pos_core[(size_t)p]; continue; }
    const int q = q_of(p);
    if (q < 0) { set_err("refresh: basis bookkeeping inconsistent (new core column without an eta)"); return MLP_INVALID; }
    rowsrc[t] = -1 - q;
  }
  for (int c = 0; c < k_new; ++c) {
    const int32_t r = R[(size_t)c];
    if (e->h_row_core[(size_t)r] >= 0) { colsrc[c] = e->h_row_core[(size_t)r]; continue; }
    int32_t p = e->h_rowcover_f[(size_t)r];  // position of the row's slack in the FACTORIZED basis: incremental_sets may have
    for (size_t q = 0; q < e->h_rc_old_rows.size(); ++q)  // moved h_rowcover_f on to the new basis already
      if (e->h_rc_old_rows[q] == r) { p = e->h_rc_old_vals[q]; break; }
    if (p < 0) { set_err("refresh: basis bookkeeping inconsistent (new core row without a basic slack)"); return MLP_INVALID; }
    colsrc[c] = -1 - p;
  }
  const int nq = (int)wrow.size();
  for (int q = 0; q < nq; ++q) wr[q] = wrow[(size_t)q];
  ST(stage_put(e, e->rf_map, map.data(), (size_t)3 * k_new + K + nq));
  const int32_t *d_rowsrc = e->rf_map, *d_colsrc = d_rowsrc + k_new, *d_jposn = d_colsrc + k_new, *d_etasrc = d_jposn + k_new,
                *d_wrow = d_etasrc + K;
  if (nq > 0)
    LAUNCH(e, k_rf_w, dim3(cdiv(k_old, 256), (unsigned)nq), 256, 0, e->dcsr_ptr, e->dcsr_idx, e->dcsr_val, d_wrow, k_old, e->Cinv, ld, e->rf_W, ld);
  LAUNCH(e, k_rf_t, cdiv(k_new, 32), 256, 0, e->Ginv, e->Kcap, K, d_etasrc, e->etaR, d_colsrc, k_new, e->Cinv, ld, e->rf_W, ld, e->rf_T);
  double* Cn = e->LUc;
  LAUNCH(e, k_rf_x0, dim3(cdiv(k_new, 256), (unsigned)std::min(k_new, 16384)), 256, 0, d_rowsrc, d_jposn, d_colsrc, k_new, e->Cinv, ld, e->rf_W, ld, Cn);
  LAUNCH(e, k_rf_ep, dim3(cdiv(k_new, 256), (unsigned)K), 256, 0, e->E, e->mld, d_jposn, k_new, e->rf_Ep, ld);
  for (int j0 = 0; j0 < K; j0 += GB_K)
    LAUNCH(e, k_gemm_sub<true>, dim3(cdiv(k_new, GB_T), cdiv(k_new, GB_T)), 256, 0, k_new, k_new, std::min(GB_K, K - j0),
           e->rf_Ep + (size_t)j0 * ld, ld, e->rf_T + j0, (int64_t)RF_CAP, Cn, ld);
  std::swap(e->Cinv, e->LUc);
  e->cnt.refreshes += 1;
  return MLP_OK;
}

// BasisSolver::reset (solver.rs:1286-1303) for B = [D | E_S], see DESIGN.md §4.  allow_refresh: the caller (mlp_pivot) has
// pushed the eta of the pivot that triggers the refactorization, so the eta file describes the whole change of the basis
// since the factors were made and may be folded into C^-1 instead of factorizing (refresh_inverse above).
static mlp_status refactor_impl(mlp_engine* e, bool allow_refresh = false) {
  const int64_t m = e->m, ng = e->ng;
  refac_stage(e, allow_refresh || e->refac_in_pivot ? "pivot: read-back, enter" : nullptr);
  e->refac_in_pivot = false;
  std::vector<int32_t> jpos, jslot, jvar, rowcover, R, prow, pval, gone;
  // Sparse storage, every basis change since the last refactorization on record: the new sets follow from the old ones.
  const bool incremental = e->sparse && e->chg_complete && (int64_t)e->h_pos_core.size() == m && (int64_t)e->h_rowcover_f.size() == m &&
                           (int64_t)e->h_Jpos_f.size() == e->k && (int64_t)e->h_R_sorted.size() == e->k;
  if (incremental) {
    ST(incremental_sets(e, jpos, R, prow, pval, gone));
    for (int32_t p : jpos) { jvar.push_back((int32_t)e->h_bvar[(size_t)p]); jslot.push_back(e->h_slot_of_row[(size_t)p]); }
  } else {
    rowcover.assign((size_t)m, -1);
    for (int64_t p = 0; p < m; ++p) {
      const int64_t v = e->h_bvar[p];
      if (v < ng) {
        jpos.push_back((int32_t)p);
        jvar.push_back((int32_t)v);
        if (!e->sparse && e->h_slot_of_row[p] < 0) { set_err("refactor: basic structural column missing from the cache"); return MLP_INVALID; }
        jslot.push_back(e->h_slot_of_row[p]);
      } else rowcover[v - ng] = (int32_t)p;
    }
    for (int64_t i = 0; i < m; ++i) if (rowcover[i] < 0) R.push_back((int32_t)i);
  }
  const int64_t k = (int64_t)jpos.size();
  if (!incremental && e->sparse && k > 1) {
    // order_simple (ordering.rs:4-21): columns by ascending entry count, FIFO — i.e. ascending basis position — within a
    // count.  (For a dense A every column has m entries and the order is the basis-position order built above.)
    std::vector<int32_t> ord((size_t)k);
    for (int64_t t = 0; t < k; ++t) ord[(size_t)t] = (int32_t)t;
    auto cnt = [&](int32_t t) { return e->h_csc_ptr[(size_t)jvar[(size_t)t] + 1] - e->h_csc_ptr[(size_t)jvar[(size_t)t]]; };
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return cnt(a) < cnt(b); });
    std::vector<int32_t> p2((size_t)k), v2((size_t)k), s2((size_t)k);
    for (int64_t t = 0; t < k; ++t) { p2[(size_t)t] = jpos[(size_t)ord[(size_t)t]]; v2[(size_t)t] = jvar[(size_t)ord[(size_t)t]]; s2[(size_t)t] = jslot[(size_t)ord[(size_t)t]]; }
    jpos.swap(p2); jvar.swap(v2); jslot.swap(s2);
  }
  if ((int64_t)R.size() != k) { set_err("refactor: basis bookkeeping inconsistent"); return MLP_INVALID; }
  std::vector<int32_t> Rsorted(R);  // R itself is overwritten with the factors' row order after a true factorization
  refac_stage(e, "host: index sets + column order");
  for (int32_t sl : e->h_pending_free) e->h_free_slots.push_back(sl);
  e->h_pending_free.clear();
  for (int l = 0; l < 2; ++l) CU(cudaStreamSynchronize(e->lane[l].st));
  e->spec_var = -1;
  e->ftran_var = -1;
  refac_stage(e, "drain both lanes");
  {
    size_t segs = 0;
    if (e->sparse) for (int32_t v : jvar) segs += (size_t)(e->h_col_seg[(size_t)v + 1] - e->h_col_seg[(size_t)v]);
    ST(stage_begin(e, 2 * (size_t)m + 10 * (size_t)k + segs + 2 * (size_t)e->K + 4 * RF_MAXK + 4 * prow.size() + 256));
  }
  bool refreshed = false;
  int64_t rf_core_before = 0;
  if (allow_refresh && k > 0 && k <= e->kcap && can_refresh(e)) {
    ST(refresh_inverse(e, jpos, R));
    refreshed = true;
    refac_stage(e, "refresh: C^-1 from the eta file");
  }
  ST(ensure_lu_capacity(e, k));
  // eta arena: the reference allows eta nnz up to lu nnz (solver.rs:1096-1097) ~ (k+1) dense columns
  {
    // The arena is dense, m doubles per eta, and bounded at 16 GB — a full arena just forces the next refactorization.  Dense
    // A: lu nnz ~ m k and an eta has m entries, so the file holds up to ~k etas: reserve 2k + 32.  Sparse A: the file is
    // short (lu nnz / nnz(alpha): tens to hundreds of etas) — reserving 2k + 32 columns would re-allocate gigabytes every
    // time k doubles (measured: 0.8 s per growth with peer mappings in place); follow the file's own length instead.
    const int64_t by_mem = std::max<int64_t>(1024, ((int64_t)16 << 30) / (8 * e->mld));
    const int64_t want = e->sparse ? std::min<int64_t>(2 * k + 32, 4 * e->K + 128) : 2 * k + 32;
    ST(ensure_eta_capacity(e, std::min<int64_t>(want, by_mem)));
  }
  refac_stage(e, "capacity (LU, eta arena)");
  e->k = k;
  e->K = 0;
  CU(cudaMemsetAsync(e->d_res->flags + 1, 0, sizeof(int), e->stream));
  CU(cudaMemsetAsync(e->etaLast, 0xff, (size_t)e->mld * sizeof(int32_t), e->stream));  // eta file is empty: no chains
  CU(cudaMemsetAsync(e->touched, 0, (size_t)e->mld, e->stream));
  // device maps by patches (the touched slack rows fit the scratch behind the refresh's maps) or in full
  const bool patch_maps = incremental && prow.size() <= (size_t)RF_MAXK && e->rf_map != nullptr;
  if (incremental) for (int32_t p : e->h_eta_pos) e->h_last_eta_of_row[(size_t)p] = -1;
  else std::fill(e->h_last_eta_of_row.begin(), e->h_last_eta_of_row.end(), -1);
  if (patch_maps) {
    int32_t* d_patch = e->rf_map + 3 * e->kcap;  // 2 RF_MAXK entries; stream-ordered behind the refresh kernels that read this area
    if (!prow.empty()) {  // rowcover: only the slack rows the basis changes touched
      ST(stage_put(e, d_patch, prow.data(), prow.size()));
      ST(stage_put(e, d_patch + RF_MAXK, pval.data(), pval.size()));
      LAUNCH(e, k_patch_i32, cdiv((int64_t)prow.size(), 256), 256, 0, e->rowcover, (const int32_t*)d_patch, (const int32_t*)(d_patch + RF_MAXK), (int)prow.size());
    }
    if (!gone.empty()) {  // rowcore: rows that left the core (a subset of the touched rows); the rows of the new core are set below
      std::vector<int32_t> minus((size_t)gone.size(), -1);
      ST(stage_put(e, d_patch, gone.data(), gone.size()));
      ST(stage_put(e, d_patch + RF_MAXK, minus.data(), minus.size()));
      LAUNCH(e, k_patch_i32, cdiv((int64_t)gone.size(), 256), 256, 0, e->rowcore, (const int32_t*)d_patch, (const int32_t*)(d_patch + RF_MAXK), (int)gone.size());
    }
  } else if (incremental) {
    ST(stage_put(e, e->rowcover, e->h_rowcover_f.data(), (size_t)m));
  } else {
    ST(stage_put(e, e->rowcover, rowcover.data(), (size_t)m));
  }
  if (e->sparse && !patch_maps) {  // also for an empty core: later refactorizations patch this map
    std::vector<int32_t> rowcore((size_t)m, -1);
    for (int64_t i = 0; i < k; ++i) rowcore[R[i]] = (int32_t)i;
    ST(stage_put(e, e->rowcore, rowcore.data(), (size_t)m));
  }
  if (k > 0) {
    ST(stage_put(e, e->Jpos, jpos.data(), (size_t)k));
    ST(stage_put(e, e->Jslot, jslot.data(), (size_t)k));
    ST(stage_put(e, e->Rp, R.data(), (size_t)k));
    if (e->sparse) {
      if (e->corevar_k > 0) LAUNCH(e, k_set_corepos, cdiv(e->corevar_k, 256), 256, 0, e->corepos, e->corevar, (int)e->corevar_k, 1);
      ST(stage_put(e, e->corevar, jvar.data(), (size_t)k));
      LAUNCH(e, k_set_corepos, cdiv(k, 256), 256, 0, e->corepos, e->corevar, (int)k, 0);
      e->corevar_k = k;
      if (patch_maps) LAUNCH(e, k_set_rowcore, cdiv(k, 256), 256, 0, e->rowcore, (const int32_t*)e->Rp, (int)k);
      // the core's segments
      std::vector<int32_t> cid, cfirst((size_t)k + 1, 0);
      for (int64_t t = 0; t < k; ++t) {
        cfirst[t] = (int32_t)cid.size();
        for (int64_t sg = e->h_col_seg[jvar[t]]; sg < e->h_col_seg[jvar[t] + 1]; ++sg) cid.push_back((int32_t)sg);
      }
      cfirst[k] = (int32_t)cid.size();
      e->ncseg = (int64_t)cid.size();
      if (e->ncseg > e->cseg_cap) {
        PoolScope pool(e->use_pool ? e->stream : nullptr);
        dev_free(e->cseg_id); dev_free(e->csum[0]); dev_free(e->csum[1]);
        e->cseg_cap = std::max<int64_t>(4 * e->ncseg, 1 << 15);
        ST(dev_alloc(&e->cseg_id, e->cseg_cap)); ST(dev_alloc(&e->csum[0], e->cseg_cap)); ST(dev_alloc(&e->csum[1], e->cseg_cap));
      }
      ST(stage_put(e, e->cseg_id, cid.data(), cid.size()));
      ST(stage_put(e, e->cseg_first, cfirst.data(), cfirst.size()));
    }
  refac_stage(e, "uploads: index maps, core segments");
    if (e->sparse) ST(build_core_rows(e, jvar));
  refac_stage(e, "compact core rows (DCSR)");
    if (refreshed) {
      // C^-1 is already the new core's.  Probe it against the new core (max |C C^-1 - I| over sampled columns) and count the
      // core's entries for the estimate of LUFactors::nnz below — one read-back; a failed probe falls through to the true
      // factorization (LUc, the old inverse's buffer, is scratch again).
      double r;
      ST(probe_inverse(e, k, &r, &rf_core_before));
      if (r <= e->rf_tol) e->rf_worst = std::max(e->rf_worst, r);
      if (!(r <= e->rf_tol)) { refreshed = false; e->cnt.refresh_rejects += 1; e->cnt.refreshes -= 1; }
  refac_stage(e, "refresh: accuracy probe + read back");
    }
    if (!refreshed) {
    if (e->sparse) {
      CU(cudaMemsetAsync(e->LUc, 0, (size_t)e->kcap * k * sizeof(double), e->stream));
      LAUNCH(e, k_extract_core_seg, cdiv(e->ncseg, 8), 256, 0, e->csc_ptr, e->csc_idx, e->csc_val, e->seg_col, e->seg_off, e->cseg_id,
             (int)e->ncseg, e->corepos, e->rowcore, e->LUc, e->kcap);
      CU(cudaMemsetAsync(e->d_nnzcnt, 0, 2 * sizeof(unsigned long long), e->stream));
      LAUNCH(e, k_core_row_counts, cdiv(k, 256), 256, 0, e->LUc, e->kcap, (int)k, e->lu_rcnt, e->d_nnzcnt);
    } else
      LAUNCH(e, k_extract_core, dim3(cdiv(k, 256), (unsigned)k), 256, 0, e->Bcols, e->mld, (int)k, e->Rp, e->Jslot, e->LUc, e->kcap);
  refac_stage(e, "extract core + row counts");
    int* flags = e->d_res->flags;
    for (int j0 = 0; j0 < (int)k;) {
      const int rows = (int)k - j0;
      // widest panel whose rows x nb block (+ row ids, permutation) fits in shared memory; else work in place in global memory
      int nb = LU_NB, use_smem = 0;
      const size_t per_row = e->sparse ? 12 : 8;  // row ids + permutation (+ row entry counts)
      for (int cand = LU_NB; cand >= 4; cand /= 2)
        if ((size_t)rows * cand * 8 + (size_t)rows * per_row <= e->smem_optin) { nb = cand; use_smem = 1; break; }
      nb = std::min(nb, rows);
      const size_t smem = use_smem ? (size_t)rows * nb * 8 + (size_t)rows * per_row : 0;
      const int pt = std::max(64, std::min(1024, (rows + 31) / 32 * 32));  // one row per thread
      LAUNCH(e, k_lu_panel, 1, pt, smem, e->LUc, e->kcap, (int)k, j0, nb, e->Rp, e->sparse ? e->lu_rcnt : (int32_t*)nullptr, flags,
             e->lu_aff, e->lu_aff + 64, e->lu_aff + 128, e->lu_perm, use_smem);
      if ((int)k > nb) LAUNCH(e, k_lu_swap_solve, cdiv(k - nb, 8), 256, 0, e->LUc, e->kcap, (int)k, j0, nb, e->lu_aff, e->lu_aff + 64,
                              e->lu_aff + 128, flags);
      const int rem = rows - nb;
      if (rem > 0) LAUNCH(e, k_lu_trailing, dim3(cdiv(rem, LU_NC), cdiv(rem, 256)), 256, 0, e->LUc, e->kcap, (int)k, j0, nb, flags);
      j0 += nb;
    }
  refac_stage(e, "LU panels / swap-solve / trailing");
    {  // (L U)^-1, one CTA per column
      const size_t need = (size_t)k * sizeof(double);
      const int use_smem = need <= e->smem_optin ? 1 : 0;
      if (k >= e->inv_blocked_min) {
        // blocked substitution on all columns at once (dense_block.cuh): X = I; forward through L, backward through U
        const int nbk = cdiv(k, 32);
        LAUNCH(e, k_set_identity, dim3(cdiv(k, 256), (unsigned)k), 256, 0, e->Cinv, e->kcap, (int)k);
        for (int b = 0; b < nbk; ++b) {  // L y = e: X stays lower triangular, only columns < (b+1)*32 are non-zero
          const int r0 = b * 32, nb = std::min<int>(32, (int)k - r0), nc = std::min<int>((int)k, r0 + nb), below = (int)k - (r0 + nb);
          LAUNCH(e, k_tri_block<true>, cdiv(nc, 128), 128, 0, e->LUc, e->kcap, r0, nb, e->Cinv, e->kcap, nc);
          if (below > 0)
            LAUNCH(e, k_gemm_sub<true>, dim3(cdiv(below, GB_T), cdiv(nc, GB_T)), 256, 0, below, nc, nb, e->LUc + (size_t)r0 * e->kcap + r0 + nb,
                   e->kcap, e->Cinv + r0, e->kcap, e->Cinv + r0 + nb, e->kcap);
        }
        for (int b = nbk - 1; b >= 0; --b) {  // U x = y over all k columns
          const int r0 = b * 32, nb = std::min<int>(32, (int)k - r0);
          LAUNCH(e, k_tri_block<false>, cdiv(k, 128), 128, 0, e->LUc, e->kcap, r0, nb, e->Cinv, e->kcap, (int)k);
          if (r0 > 0)
            LAUNCH(e, k_gemm_sub<true>, dim3(cdiv(r0, GB_T), cdiv(k, GB_T)), 256, 0, r0, (int)k, nb, e->LUc + (size_t)r0 * e->kcap, e->kcap,
                   e->Cinv + r0, e->kcap, e->Cinv, e->kcap);
        }
      } else if (k <= 256) LAUNCH(e, k_core_inverse_pf<1>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
      else if (k <= 512) LAUNCH(e, k_core_inverse_pf<2>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
      else if (k <= 1024) LAUNCH(e, k_core_inverse_pf<4>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
      else if (k <= 2048) LAUNCH(e, k_core_inverse_pf<8>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
      else if (k <= 4096) LAUNCH(e, k_core_inverse_pf<16>, (unsigned)k, 256, need, e->LUc, e->kcap, (int)k, e->Cinv, flags);
      else LAUNCH(e, k_core_inverse, (unsigned)k, 256, use_smem ? need : 0, e->LUc, e->kcap, (int)k, e->Cinv, flags, use_smem);
    }
  refac_stage(e, "explicit inverse");
    if (e->sparse) {
      LAUNCH(e, k_count_offdiag, dim3(cdiv(k, 256), cdiv(k, 64)), 256, 0, e->LUc, e->kcap, (int)k, e->d_nnzcnt + 1);
      CU(cudaMemcpyAsync(&e->d_res->i[4], e->d_nnzcnt, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, e->stream));
    }
    ST(fetch_res(e, e->lane[0]));
    if (e->h_res->flags[1]) { set_err("singular basis"); return MLP_SINGULAR; }
    // the factorization permuted the core's rows (Rp): column c of C^-1 belongs to row Rp[c] — the next refresh needs that order
    if (e->sparse) ST(d2h(e, R.data(), e->Rp, (size_t)k * sizeof(int32_t)));
    if (e->sparse && e->refac_trace) {  // calibration of the refresh probe: the same measure on a freshly factorized inverse
      double r;
      int64_t ce;
      ST(probe_inverse(e, k, &r, &ce));
      e->rf_worst_true = std::max(e->rf_worst_true, r);
    }
    }
  refac_stage(e, "count off-diagonal + read back");
  } else {
    if (e->sparse) ST(build_core_rows(e, jvar));  // empty
    if (e->sparse && e->corevar_k > 0) {
      LAUNCH(e, k_set_corepos, cdiv(e->corevar_k, 256), 256, 0, e->corepos, e->corevar, (int)e->corevar_k, 1);
      e->corevar_k = 0;
    }
    CU(cudaStreamSynchronize(e->stream));
  }
  // LUFactors::nnz (lu.rs:52-54): lower.nondiag + upper.nondiag + m.  The entries of the k structural basic columns in
  // slack-covered rows go to U unchanged; the core contributes the off-diagonal entries of its factors, FILL-IN INCLUDED
  // (counted on the device; exact zeros are not stored, lu.rs:253-255).  Dense A: no zeros, k(k-1) + (m-k)k + m in closed
  // form.  The refactor rule (solver.rs:1096-1097) is the reference's, applied to the factors the engine really has: with
  // the same column order and pivot rule (ties aside) their size tracks the reference's.
  if (e->sparse) {
    int64_t nz = 0;
    for (int32_t v : jvar) nz += e->h_csc_ptr[(size_t)v + 1] - e->h_csc_ptr[(size_t)v];
    if (refreshed) {
      // no factors to count: the part outside the core is exact, the core's L\U is taken to fill as it did at the last true
      // factorization (off-diagonal entries of the factors per entry of the core)
      e->lu_nnz = (nz - rf_core_before) + (int64_t)((double)rf_core_before * e->fill_true) + m;
    } else {
      const int64_t core_before = k > 0 ? e->h_res->i[4] : 0, core_offdiag = k > 0 ? e->h_res->i[5] : 0;
      e->lu_nnz = (nz - core_before) + core_offdiag + m;
      e->fill_true = core_before > 0 ? (double)core_offdiag / (double)core_before : 1.0;
      e->pivots_since_lu = 0;
    }
  } else e->lu_nnz = k * (k - 1) + (m - k) * k + m;
  if (e->sparse) {  // the sets of the factorized basis, for the next refresh / the next incremental set-up
    if (incremental) {
      for (int32_t p : e->h_Jpos_f) e->h_pos_core[(size_t)p] = -1;
      for (int32_t r : e->h_R_f) e->h_row_core[(size_t)r] = -1;
    } else {
      e->h_pos_core.assign((size_t)m, -1);
      e->h_row_core.assign((size_t)m, -1);
      e->h_rowcover_f.swap(rowcover);
    }
    for (int64_t t = 0; t < k; ++t) { e->h_pos_core[(size_t)jpos[(size_t)t]] = (int32_t)t; e->h_row_core[(size_t)R[(size_t)t]] = (int32_t)t; }
    e->h_Jpos_f.swap(jpos);
    e->h_R_f.swap(R);          // the factors' row order (permuted by a true factorization)
    e->h_R_sorted.swap(Rsorted);
    e->h_eta_pos.clear();
    e->h_eta_leave.clear();
    e->h_rc_old_rows.clear();
    e->h_rc_old_vals.clear();
    e->chg_complete = true;
    e->inv_valid = true;
  }
  ST(stage_end(e));
  e->cnt.refactors += 1;
  e->cnt.k_structural = k;
  e->refac_k_sum += (double)k;
  refac_stage(e, "host: lu nnz");
  return MLP_OK;
}

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

mlp_status mlp_btran_unit(mlp_engine* e, int64_t row) {
  if (!e || !e->initialized || row < 0 || row >= e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  Lane& l1 = e->lane[1];
  ST(begin1(e));
  // c = e_row and, with etas, u = row `row` of E (solver.rs:1326-1330 for a unit vector) in one launch
  if (e->merge_small && e->K > 0 && e->K <= FE_MAXK) {  // short eta file: ... and s = (I+G)^-T u as well
    LAUNCHS(e, l1.st, k_unit_eta_t, cdiv(std::max<int64_t>(e->m, 32 * e->K), 256), 256, 0, e->work_mb, e->m, row, e->E, e->mld, e->Ginv, e->Kcap,
            (int)e->K, l1.tK2);
    ST(btran(e, l1, e->work_mb, (int)row, e->rho, true, true));
  } else {
    LAUNCHS(e, l1.st, k_unit_and_gather, cdiv(std::max<int64_t>(e->m, e->K), 256), 256, 0, e->work_mb, e->m, row, e->E, e->mld, (int)e->K, l1.tK);
    ST(btran(e, l1, e->work_mb, (int)row, e->rho, true));
  }
  // inv_basis_row_coeffs as a sparse list + |rho|^2 (solver.rs:683, 1160)
  compact(e, l1, e->rho, e->list_idx, e->list_val, e->icnt, e->scal + 1);
  ST(mark1(e));
  return MLP_OK;
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

mlp_status mlp_select_row_dual(mlp_engine* e, mlp_dual_row* out) {
  if (!e || !e->initialized) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  const int grid = std::min(cdiv(e->m, 256), 1024);
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  LAUNCH(e, k_select_row_dual, grid, 256, 0, e->xB, e->loB, e->hiB, e->w, (int)e->m, e->enable_dse, l0.red_f, l0.red_i,
         l0.red_counter, e->d_res);
  ST(mark0(e));
  ST(fetch_res(e, l0));
  out->row = e->h_res->i[0];
  out->val = e->h_res->f[0];
  out->min = e->h_res->f[1];
  out->max = e->h_res->f[2];
  e->dual_row_host = out->row;
  e->dual_row_val = out->val;
  return MLP_OK;
}

mlp_status mlp_ratio_dual(mlp_engine* e, int64_t row, double leaving_new_val, mlp_dual_entering* out) {
  if (!e || !e->initialized || row < 0 || row >= e->m) return MLP_INVALID;
  CU(cudaSetDevice(e->device));
  // leaving_diff_sign = leaving_new_val > basic_var_vals[row] (solver.rs:925)
  Lane& l0 = e->lane[0];
  ST(begin0(e));
  e->sel_valid = false;  // the candidate buffer is reused for the dual ratio test
  double bv = e->dual_row_val;  // basic_var_vals[row] as choose_pivot_row_dual saw it; x_B has not changed since
  if (e->dual_row_host != row) ST(d2h(e, &bv, e->xB + row, sizeof(double)));
  const int lds = leaving_new_val > bv ? 1 : 0;
  const int grid = std::min(cdiv(e->nt, 256), 1024);
  LAUNCH(e, k_ratio_dual_1, grid, 256, 0, e->rc, e->d, e->vflag, e->nt, lds, l0.red_f, l0.red_counter, e->scal);
  if (e->world > 1) {  // Harris pass 1 is a min over ALL variables: all-gather the shard minima
    ST(e->comm->allgather(e->scal, e->xred, sizeof(double), e->stream));
    LAUNCH(e, k_min_small, 1, 1, 0, e->xred, e->world, e->scal);
  }
  LAUNCH(e, k_ratio_dual_2, grid, 256, 0, e->rc, e->d, e->vflag, e->vpos, e->xnb, e->nt, e->n, e->c0, e->ng, lds, e->scal,
         l0.red_f, l0.red_i, l0.red_counter, e->d_res->flags, (Cand*)e->xsend, e->rank == 0 ? 1 : 0);
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
