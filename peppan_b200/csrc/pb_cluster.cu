// Gene clustering: pb_cluster.  Stands in for what getClust obtains from
// `mmseqs createdb / linclust --min-seq-id I -c C / createtsv` (modules/clust.py:62-66) together
// with the outcome rule getClust imposes on top of it: the representative of a cluster is its
// first member in input (priority) order (modules/clust.py:72-85).
//
// Definition (restated by the scalar greedy of the CPU oracle): genes are visited in input
// order; gene b joins the EARLIEST representative a < b that has a verified edge to it, else b
// becomes a representative.  An edge (a, b) is verified when the local alignment found by the
// nucleotide search (coding strand only) has identity >= I and covers >= C of both genes
// (MMseqs2 --cov-mode 0).  Because only representatives can be joined, genes are processed in
// blocks, each in two phases: the block is searched against the representatives so far (a gene
// with a verified edge joins the earliest such representative -- all of them precede the block),
// and only the genes nobody claimed are compared all against all.  The work is linear in the
// number of genes for redundant inputs, like linclust, and the result is identical to the
// all-vs-all greedy.
//
// The greedy assignment itself runs on the device as a monotone fixed-point iteration: a gene is
// decided once all its earlier neighbours are decided (K3).
#include "pb_common.h"
#include "pb_memo.h"
#include <algorithm>
#include <thread>
#include <vector>

namespace {

enum : int { UNDECIDED = 0, REP = 1, MEMBER = 2 };

// nodes: genes of the current block (local index x -> global first + x).  adj: CSR of earlier neighbours (global ids,
// ascending).  Neighbours < first are representatives by construction.
__global__ void greedy_round_kernel(const int* adj_off, const int* adj, int nb, int first, int* state, int* rep_of, int* n_changed)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nb || state[x] != UNDECIDED) return;
    int decided = REP, rep = first + x;
    for (int e = adj_off[x]; e < adj_off[x + 1]; ++e) {
        const int a = adj[e];
        int sa = a < first ? REP : state[a - first];
        if (sa == REP) { decided = MEMBER; rep = a; break; }
        if (sa == UNDECIDED) { decided = UNDECIDED; break; }
    }
    if (decided != UNDECIDED) {
        rep_of[x] = rep;
        __threadfence();
        state[x] = decided;
        atomicAdd(n_changed, 1);
    }
}

}  // namespace

extern "C" int pb_cluster(pb_ctx* ctx, const pb_seqset* genes, float min_id, float min_cov, int32_t* rep_of, pb_cluster_stats* stats)
{
    return pb_cluster_ex(ctx, genes, min_id, min_cov, 0, 11, rep_of, stats);
}

extern "C" int pb_cluster_ex(pb_ctx* ctx, const pb_seqset* genes, float min_id, float min_cov, int translate, int gtable,
                             int32_t* rep_of, pb_cluster_stats* stats)
{
    if (!ctx || !genes || !rep_of || genes->n < 0) { pb_set_error(ctx, "pb_cluster: invalid argument"); return PB_ERR_ARG; }
    pb_cluster_stats st; memset(&st, 0, sizeof(st));
    const int64_t n = genes->n;
    if (n == 0) { if (stats) *stats = st; return PB_OK; }
    if (n > 0x7fffffff) { pb_set_error(ctx, "pb_cluster: too many genes for one call"); return PB_ERR_LIMIT; }
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t sm = ctx->stream;
    PB_CUDA(ctx, cudaEventRecord(ctx->ev[13], sm));
    int64_t BLOCK_RES = 24ll << 20;                // residues of new genes per block (PB_CLUSTER_BLOCK overrides: test aid)
    if (const char* e = getenv("PB_CLUSTER_BLOCK")) { long long v = atoll(e); if (v > 0) BLOCK_RES = v; }
    // content hash of every gene: the key under which the alignments of its pairs are remembered across calls (pb_memo.h)
    std::vector<uint64_t> ghash((size_t)n);
    {
        const int nth = (int)std::min<int64_t>(8, std::max<int64_t>(1, n / 4096));
        std::vector<std::thread> th;
        for (int k = 0; k < nth; ++k)
            th.emplace_back([&, k]() { for (int64_t i = k; i < n; i += nth) ghash[i] = pb_seq_hash(genes->residues + genes->offsets[i], genes->offsets[i + 1] - genes->offsets[i]) ^ (translate ? 0x5bd1e995ull : 0ull); });
        for (auto& t : th) t.join();
    }
    int64_t memo_stats[2] = {0, 0};
    std::vector<uint64_t> rep_hash, qhash, thash;
    std::vector<int> reps;                          // global ids of the representatives so far
    std::vector<uint8_t> tbuf; std::vector<int64_t> toff;
    std::vector<uint8_t> rep_bytes; std::vector<int64_t> rep_off(1, 0);
    int64_t first = 0;
    while (first < n) {
        int64_t last = first, res = 0;
        while (last < n && (last == first || res + (genes->offsets[last + 1] - genes->offsets[last]) <= BLOCK_RES)) {
            res += genes->offsets[last + 1] - genes->offsets[last]; ++last;
        }
        const int nb = (int)(last - first), nr = (int)reps.size();
        const uint8_t* bsrc = genes->residues + genes->offsets[first];
        std::vector<int64_t> qoff(nb + 1);
        for (int i = 0; i <= nb; ++i) qoff[i] = genes->offsets[first + i] - genes->offsets[first];
        pb_search_params prm; memset(&prm, 0, sizeof(prm));
        // translate: genes compared as proteins, frame 1 against frame 1 (clust -a translates frame 1 only, modules/clust.py:42);
        // reserved[0] bit 1 pins the query frame to 1, hits on target frames 2 / 3 are ignored below
        prm.mode = translate ? PB_MODE_PROT3_SELF : PB_MODE_NT; prm.gtable = gtable; prm.min_id = min_id - 0.005f; prm.min_cov = 0;
        prm.min_ratio = std::max(0.f, min_cov - 0.005f);
        // no per-query cap: the cap of pb_search is by raw score over ALL targets, applied before later genes / members are
        // discarded here, so in a family of > 1000 genes it could cut the edge to the earliest representative
        prm.max_hits_per_query = 0x7fffffff; prm.reserved[0] = translate ? 2 : 1;
        auto verified = [&](const pb_hits& hits, const pb_hit& x) {
            if (translate && x.frame != 1) return false;
            int gapb = 0;
            if (x.cigar_n == 0) gapb = (int)x.cigar_off * (translate ? 3 : 1);       // memo path: gap bases instead of a CIGAR (pb_search_memo)
            for (uint32_t k = 0; k < x.cigar_n; ++k) { uint32_t op = hits.cigar[x.cigar_off + k]; if (op & 3) gapb += (int)(op >> 2); }
            const int nm = x.aln_len - x.mismatch - gapb;
            const double iden = (double)nm / (double)x.aln_len;
            const double qc = (double)(x.q_end - x.q_start + 1) / (double)x.q_len, sc = (double)(x.s_end - x.s_start + 1) / (double)x.s_len;
            return iden + 1e-9 >= (double)min_id && qc + 1e-9 >= (double)min_cov && sc + 1e-9 >= (double)min_cov;
        };
        // Multi-GPU (ctx->world > 1, every rank calls with the same genes): the queries of both phases are dealt round-robin
        // over the ranks (genes arrive longest first, so the deal is balanced), every rank searches its share against the
        // replicated targets, and the per-gene results are exchanged over NCCL: an element-wise maximum of joined[] after
        // phase 1, an allgather of the verified edges after phase 2.  The greedy then runs identically on every rank.
        const int W = ctx->world, R = ctx->rank;
        auto my_share = [&](const std::vector<int>& idx_all, std::vector<uint8_t>& bytes, std::vector<int64_t>& off, std::vector<int>& mine) {
            bytes.clear(); off.assign(1, 0); mine.clear();
            for (size_t k = (size_t)R; k < idx_all.size(); k += (size_t)W) {
                const int i = idx_all[k];
                bytes.insert(bytes.end(), bsrc + qoff[i], bsrc + qoff[i + 1]);
                off.push_back(off.back() + (qoff[i + 1] - qoff[i])); mine.push_back(i);
            }
        };
        std::vector<uint8_t> sbytes; std::vector<int64_t> soff; std::vector<int> smine;
        // phase 1: the block against the representatives so far.  Every representative precedes every gene of the block, so
        // a gene with a verified edge to one joins the earliest such representative whatever happens inside the block.
        std::vector<int> joined(nb, -1);
        if (nr > 0) {
            pb_seqset qs{bsrc, qoff.data(), nb}, ts{rep_bytes.data(), rep_off.data(), nr};
            if (W > 1) {
                std::vector<int> everyone(nb);
                for (int i = 0; i < nb; ++i) everyone[i] = i;
                my_share(everyone, sbytes, soff, smine);
                qs = pb_seqset{sbytes.data(), soff.data(), (int64_t)smine.size()};
            }
            qhash.clear();
            if (W > 1) for (int i : smine) qhash.push_back(ghash[first + i]);
            else for (int i = 0; i < nb; ++i) qhash.push_back(ghash[first + i]);
            pb_hits hits; pb_search_stats sst;
            int rc = pb_search_memo(ctx, &qs, &ts, &prm, &hits, &sst, qhash.data(), rep_hash.data(), memo_stats);
            if (rc) return rc;
            st.n_pairs_verified += sst.n_windows; st.sw_cells += sst.sw_cells; st.kernel_launches += sst.kernel_launches;
            for (int64_t h = 0; h < hits.n_hits; ++h) {
                const pb_hit& x = hits.hits[h];
                if (!verified(hits, x)) continue;
                ++st.n_edges;
                const int a = reps[x.s_id], b = W > 1 ? smine[x.q_id] : x.q_id;
                if (joined[b] < 0 || a < joined[b]) joined[b] = a;
            }
            pb_free_hits(&hits);
            if (W > 1) {
                // every gene was searched by exactly one rank: others hold -1; the maximum of (INT_MAX - a) keeps the earliest
                std::vector<int32_t> v(nb);
                for (int i = 0; i < nb; ++i) v[i] = joined[i] < 0 ? -1 : 0x7fffffff - joined[i];
                rc = pb_allreduce_max_i32(ctx, v.data(), nb); if (rc) return rc;
                for (int i = 0; i < nb; ++i) joined[i] = v[i] < 0 ? -1 : 0x7fffffff - v[i];
            }
        }
        // phase 2: the genes no representative claimed, all against all; greedy fixed point on the device
        std::vector<int> novel;
        for (int i = 0; i < nb; ++i) { if (joined[i] >= 0) rep_of[first + i] = joined[i]; else novel.push_back(i); }
        const int nn = (int)novel.size();
        if (nn > 0) {
            tbuf.clear(); toff.assign(1, 0);
            for (int i : novel) {
                tbuf.insert(tbuf.end(), bsrc + qoff[i], bsrc + qoff[i + 1]);
                toff.push_back(toff.back() + (qoff[i + 1] - qoff[i]));
            }
            pb_seqset ns{tbuf.data(), toff.data(), nn}, nq = ns;
            // only an EARLIER gene can claim a later one: windows whose target does not precede the query are never aligned
            pb_search_params prm2 = prm;
            prm2.reserved[0] |= 4; prm2.reserved[1] = W; prm2.reserved[2] = R;
            if (W > 1) {
                std::vector<int> everyone(nn);
                for (int i = 0; i < nn; ++i) everyone[i] = novel[i];
                my_share(everyone, sbytes, soff, smine);
                nq = pb_seqset{sbytes.data(), soff.data(), (int64_t)smine.size()};
            }
            thash.clear(); qhash.clear();
            for (int i : novel) thash.push_back(ghash[first + i]);
            if (W > 1) for (int i : smine) qhash.push_back(ghash[first + i]); else qhash = thash;
            pb_hits hits; pb_search_stats sst;
            int rc = pb_search_memo(ctx, &nq, &ns, &prm2, &hits, &sst, qhash.data(), thash.data(), memo_stats);
            if (rc) return rc;
            st.n_pairs_verified += sst.n_windows; st.sw_cells += sst.sw_cells; st.kernel_launches += sst.kernel_launches;
            std::vector<int32_t> mine_edges;             // (b, a) in novel-local indices, a earlier than b, flattened
            for (int64_t h = 0; h < hits.n_hits; ++h) {
                const pb_hit& x = hits.hits[h];
                const int b = R + x.q_id * W;            // novel-local index of the query (round-robin deal)
                if (x.s_id >= b) continue;
                if (verified(hits, x)) { mine_edges.push_back(b); mine_edges.push_back(x.s_id); }
            }
            pb_free_hits(&hits);
            std::vector<int32_t> all_edges;
            rc = pb_allgather_i32(ctx, mine_edges, all_edges); if (rc) return rc;
            std::vector<std::pair<int, int>> edges(all_edges.size() / 2);
            for (size_t e = 0; e < edges.size(); ++e) edges[e] = std::make_pair(all_edges[2 * e], all_edges[2 * e + 1]);
            std::sort(edges.begin(), edges.end());
            edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
            st.n_edges += (int64_t)edges.size();
            std::vector<int> adj_off(nn + 1, 0), adj(edges.size());
            for (auto& e : edges) adj_off[e.first + 1]++;
            for (int i = 0; i < nn; ++i) adj_off[i + 1] += adj_off[i];
            for (size_t i = 0; i < edges.size(); ++i) adj[i] = edges[i].second;
            DevBuf d_off, d_adj, d_state, d_rep, d_cnt;
            PB_CUDA(ctx, d_off.alloc((nn + 1) * 4, sm)); PB_CUDA(ctx, d_adj.alloc(std::max<size_t>(adj.size(), 1) * 4, sm));
            PB_CUDA(ctx, d_state.alloc(nn * 4, sm)); PB_CUDA(ctx, d_rep.alloc(nn * 4, sm)); PB_CUDA(ctx, d_cnt.alloc(4, sm));
            PB_CUDA(ctx, cudaMemcpyAsync(d_off.p, adj_off.data(), (nn + 1) * 4, cudaMemcpyHostToDevice, sm));
            if (!adj.empty()) PB_CUDA(ctx, cudaMemcpyAsync(d_adj.p, adj.data(), adj.size() * 4, cudaMemcpyHostToDevice, sm));
            PB_CUDA(ctx, cudaMemsetAsync(d_state.p, 0, nn * 4, sm));
            int decided = 0;
            while (decided < nn) {
                PB_CUDA(ctx, cudaMemsetAsync(d_cnt.p, 0, 4, sm));
                greedy_round_kernel<<<(nn + 255) / 256, 256, 0, sm>>>(d_off.as<int>(), d_adj.as<int>(), nn, 0, d_state.as<int>(), d_rep.as<int>(), d_cnt.as<int>());
                PB_CUDA(ctx, cudaGetLastError());
                int c = 0;
                PB_CUDA(ctx, cudaMemcpyAsync(&c, d_cnt.p, 4, cudaMemcpyDeviceToHost, sm));
                PB_CUDA(ctx, cudaStreamSynchronize(sm));
                if (c == 0) { pb_set_error(ctx, "pb_cluster: greedy iteration made no progress"); return PB_ERR_LIMIT; }
                decided += c; st.greedy_rounds++; st.kernel_launches++;
            }
            std::vector<int> lrep(nn);
            PB_CUDA(ctx, cudaMemcpyAsync(lrep.data(), d_rep.p, nn * 4, cudaMemcpyDeviceToHost, sm));
            PB_CUDA(ctx, cudaStreamSynchronize(sm));
            for (int x = 0; x < nn; ++x) rep_of[first + novel[x]] = (int)first + novel[lrep[x]];
        }
        for (int i = 0; i < nb; ++i)
            if (rep_of[first + i] == first + i) {
                reps.push_back((int)(first + i)); rep_hash.push_back(ghash[first + i]);
                const int64_t a = genes->offsets[first + i], L = genes->offsets[first + i + 1] - a;
                rep_bytes.insert(rep_bytes.end(), genes->residues + a, genes->residues + a + L);
                rep_off.push_back(rep_off.back() + L);
            }
        st.n_blocks++;
        first = last;
    }
    st.n_reps = (int64_t)reps.size(); st.n_pairs_remembered = memo_stats[0];
    PB_CUDA(ctx, cudaEventRecord(ctx->ev[14], sm));
    PB_CUDA(ctx, cudaEventSynchronize(ctx->ev[14]));
    cudaEventElapsedTime(&st.ms_total, ctx->ev[13], ctx->ev[14]);
    if (stats) *stats = st;
    return PB_OK;
}
