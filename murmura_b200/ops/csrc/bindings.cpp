// murmura_b200 — Python bindings of the sm_100a extension.
#include <torch/extension.h>
#include <pybind11/stl.h>

namespace py = pybind11;
using torch::Tensor;

void bind_arena(py::module_& m);

// aggregate.cu
void publish(Tensor live, int64_t pub_ptr, int64_t stride, int64_t V, int64_t Pf, int64_t Pf_pad, c10::optional<Tensor> ints,
             Tensor scale, Tensor noise_std, Tensor node_gid, int64_t seed, int64_t round, int64_t peer_flags_ptr,
             int64_t G, int64_t my_rank, int64_t epoch, Tensor ticket);
void weighted_gather(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr,
                     Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, int64_t len, bool renorm,
                     int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr, bool use_tma);
void publish_sum(Tensor live, int64_t pub_ptr, int64_t rsum_ptr, int64_t stride, int64_t V, int64_t Pf, int64_t Pf_pad,
                 c10::optional<Tensor> ints, Tensor scale, Tensor noise_std, Tensor node_gid, int64_t seed, int64_t round,
                 int64_t peer_flags_ptr, int64_t G, int64_t my_rank, int64_t epoch, Tensor ticket);
void fedavg_fullmesh(Tensor live, int64_t pub_local_ptr, int64_t peer_rsum_tbl, int64_t mc_rsum_ptr, int64_t stride, int64_t V,
                     int64_t len, int64_t N, int64_t G, Tensor byz, Tensor rank_nodes, int64_t timed_out_ptr, int64_t tot_local_ptr);
void fullmesh_reduce_scatter(Tensor anchor, int64_t peer_rsum_tbl, int64_t mc_rsum_ptr, int64_t peer_tot_tbl, int64_t mc_tot_ptr, int64_t len,
                             int64_t G, int64_t my_rank, int64_t timed_out_ptr, int64_t peer_flags_ptr, int64_t epoch, Tensor ticket);
void nvls_fedavg(Tensor live, int64_t pub_local_ptr, int64_t mc_pub_ptr, int64_t stride, int64_t V, int64_t S, int64_t len,
                 int64_t N, Tensor byz, int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr);
void wait_epoch(Tensor anchor, int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr);
void tail_blend(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                Tensor src_slot, Tensor mask, Tensor w_tail, int64_t Pf_pad, Tensor ints, int64_t timed_out_ptr);
void edge_distances(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                    Tensor src_slot, Tensor mask, int64_t len, Tensor d2, Tensor n2, int64_t flags_ptr, int64_t G, int64_t epoch,
                    double timeout_ms, int64_t timed_out_ptr);
void pairwise_distances(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                        Tensor src_slot, Tensor mask, int64_t len, Tensor D, int64_t max_m, int64_t flags_ptr, int64_t G, int64_t epoch,
                        double timeout_ms, int64_t timed_out_ptr);
void krum_refine(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                 Tensor src_slot, Tensor mask, int64_t len, Tensor D, Tensor norms, double tau, Tensor scratch, c10::optional<Tensor> nref);
void count_sketch(int64_t base_ptr, int64_t stride, Tensor slots, Tensor table, int64_t Pf, int64_t K, Tensor out);
void sketch_quant_mxfp8(Tensor sk, int64_t q_ptr, int64_t sc_ptr, int64_t Kpad);
void fedavg_weights(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats);
void balance_filter(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                    Tensor d2, Tensor n2, Tensor dist_out, double factor, double alpha, int64_t min_neighbors, int64_t timed_out_ptr);
void sketchguard_filter(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                        Tensor own_sketch, int64_t peer_sketch_ptr, int64_t peer_q_ptr, int64_t peer_sc_ptr, int64_t plane_slots,
                        int64_t K, int64_t Kpad, bool fp8, double factor, double alpha, int64_t min_neighbors, Tensor hist,
                        Tensor dist_out, int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr);
void ubar_stage1(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                 Tensor d2, double rho, int64_t min_neighbors, Tensor cand, Tensor rank, int64_t timed_out_ptr);
void ubar_stage2(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                 Tensor cand, Tensor rank, Tensor loss, Tensor own_loss, double alpha, bool use_loss);
void krum_select(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                 Tensor D, int64_t num_compromised, Tensor winner, int64_t timed_out_ptr);
void trust_filter(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                  Tensor vac, Tensor acc, Tensor src_gid, int64_t N, Tensor ema, Tensor ema_valid, double accuracy_weight,
                  double vacuity_threshold, double momentum, bool adaptive, double threshold, double self_weight, Tensor trust_out,
                  int64_t timed_out_ptr);
// train.cu
void sgd_step(Tensor live, int64_t stride, Tensor grad, int64_t gstride, int64_t slot0, int64_t nslots, int64_t n, double lr);
void sgd_multi(std::vector<Tensor> params, std::vector<Tensor> grads, double lr);
Tensor bn_eval(Tensor x, Tensor mean, Tensor var, c10::optional<Tensor> gamma, c10::optional<Tensor> beta, double eps, bool relu,
               c10::optional<Tensor> residual);
std::vector<Tensor> ce_loss_fwd_bwd(Tensor logits, Tensor targets, c10::optional<Tensor> loss_acc);
std::vector<Tensor> gather_batch(Tensor X, Tensor y, Tensor perm, Tensor step, Tensor ticket, int64_t eb);
std::vector<Tensor> bn_act_fwd(Tensor x, c10::optional<Tensor> residual, Tensor gamma, Tensor beta, c10::optional<Tensor> rmean,
                               c10::optional<Tensor> rvar, c10::optional<Tensor> nbt, double momentum, double eps, bool relu);
std::vector<Tensor> bn_act_bwd(Tensor dy, Tensor x, Tensor y, Tensor gamma, Tensor save_mean, Tensor save_invstd, bool relu, bool want_dres);
void ce_eval(Tensor logits, Tensor targets, c10::optional<Tensor> n_valid, Tensor stats);
void dirichlet_eval(Tensor alpha, Tensor targets, c10::optional<Tensor> n_valid, Tensor stats);
std::vector<Tensor> evidential_loss_fwd_bwd(Tensor alpha, Tensor targets, double lam, c10::optional<Tensor> lam_t);
// gram_tcgen05.cu
Tensor gram_make_maps(std::vector<int64_t> base_ptrs, int64_t rows, int64_t row_stride, int64_t row_len, int64_t kb_per_stage);
int64_t gram_kb_per_stage(int64_t ngroups);
void gram_tf32(Tensor maps_cpu, std::vector<int64_t> group_map, std::vector<int64_t> group_y, int64_t kb0, int64_t kb1,
               int64_t R, Tensor out, bool zero_out, int64_t max_ctas);
// mlp_tcgen05.cu
void grouped_linear_tf32(Tensor groups, int64_t G, int64_t max_m, int64_t K, int64_t N, int64_t ldx, int64_t ldy, int64_t act, double eps);
void grouped_eval(Tensor groups, int64_t G, int64_t max_m, int64_t C, int64_t ld, bool dirichlet, Tensor stats);
// dmtt.cu
void mobility_adjacency(Tensor pos, int64_t round, double area, double range, bool ensure_connected, Tensor adj);
void liar_claims(Tensor adj, Tensor is_liar, Tensor claims);
void dmtt_update(Tensor adj, Tensor claims, Tensor collab, Tensor received, Tensor model_score, Tensor score_valid, Tensor c_hat,
                 Tensor alpha, Tensor beta, Tensor next_collab, Tensor q_out, double rho, double lam, double w_d, double w_x,
                 double tau_U, double eta, double l1, double l2, double l3, int64_t B, Tensor gids);

// conv_tcgen05.cu / layers.cu (plans are dicts built by murmura_b200/ops/conv_plan.py and parallel/fused_trainer.py)
int64_t conv_gemm(py::dict plan);
int64_t conv_tma(py::dict plan);
py::bytes tma_encode(int64_t ptr, std::vector<int64_t> dims, std::vector<int64_t> strides_bytes, std::vector<int64_t> box,
                     std::vector<int64_t> elem_strides, int64_t swizzle);
void set_pdl(bool on);
bool get_pdl();
void gather_grouped(py::dict d);
void im2col_pack(py::dict d);
void bn_fwd_grouped(py::dict d);
void bn_bwd_grouped(py::dict d);
void maxpool_fwd_grouped(py::dict d);
void maxpool_bwd_grouped(py::dict d);
void avgpool_grouped(py::dict d);
void dropout_grouped(py::dict d);
void loss_grouped(py::dict d);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "murmura_b200 sm_100a kernels";
    bind_arena(m);
    m.def("publish", &publish);
    m.def("weighted_gather", &weighted_gather, py::arg("live"), py::arg("peer_pub_ptr"), py::arg("parity_off"), py::arg("stride"), py::arg("V"),
          py::arg("row_ptr"), py::arg("src_rank"), py::arg("src_slot"), py::arg("mask"), py::arg("w"), py::arg("len"), py::arg("renorm"),
          py::arg("flags_ptr"), py::arg("G"), py::arg("epoch"), py::arg("timeout_ms"), py::arg("timed_out_ptr"), py::arg("use_tma") = false);
    m.def("tail_blend", &tail_blend);
    m.def("wait_epoch", &wait_epoch);
    m.def("nvls_fedavg", &nvls_fedavg);
    m.def("publish_sum", &publish_sum, "publish + per-rank column sum of the published rows (full-mesh FedAvg)");
    m.def("fullmesh_reduce_scatter", &fullmesh_reduce_scatter, "two-shot all-reduce of the per-rank sum rows over peer / multicast memory (one fused kernel)");
    m.def("fedavg_fullmesh", &fedavg_fullmesh, "full-mesh FedAvg from the per-rank sums (NVLS multimem.ld_reduce or peer loads)");
    m.def("edge_distances", &edge_distances);
    m.def("pairwise_distances", &pairwise_distances);
    m.def("krum_refine", &krum_refine, "exact fp32 recomputation of the cancellation-prone pairs of a Gram-derived distance table");
    m.def("count_sketch", &count_sketch);
    m.def("sketch_quant_mxfp8", &sketch_quant_mxfp8);
    m.def("fedavg_weights", &fedavg_weights);
    m.def("balance_filter", &balance_filter);
    m.def("sketchguard_filter", &sketchguard_filter);
    m.def("ubar_stage1", &ubar_stage1);
    m.def("ubar_stage2", &ubar_stage2);
    m.def("krum_select", &krum_select);
    m.def("trust_filter", &trust_filter);
    m.def("sgd_step", &sgd_step);
    m.def("sgd_multi", &sgd_multi);
    m.def("bn_eval", &bn_eval, py::arg("x"), py::arg("mean"), py::arg("var"), py::arg("gamma"), py::arg("beta"), py::arg("eps"), py::arg("relu"),
          py::arg("residual") = py::none());
    m.def("bn_act_fwd", &bn_act_fwd, "fused training BatchNorm (+residual) (+ReLU), cluster/DSMEM reduction");
    m.def("bn_act_bwd", &bn_act_bwd);
    m.def("ce_loss_fwd_bwd", &ce_loss_fwd_bwd, "fused softmax cross-entropy forward+backward (+ running-loss accumulator)");
    m.def("gather_batch", &gather_batch, "mini-batch gather through a device-side permutation; advances the step counter");
    m.def("ce_eval", &ce_eval);
    m.def("dirichlet_eval", &dirichlet_eval);
    m.def("evidential_loss_fwd_bwd", &evidential_loss_fwd_bwd);
    m.def("gram_tf32", &gram_tf32);
    m.def("gram_make_maps", &gram_make_maps);
    m.def("gram_kb_per_stage", &gram_kb_per_stage);
    m.def("grouped_linear_tf32", &grouped_linear_tf32);
    m.def("grouped_eval", &grouped_eval);
    m.def("mobility_adjacency", &mobility_adjacency);
    m.def("liar_claims", &liar_claims);
    m.def("dmtt_update", &dmtt_update);
    m.def("conv_gemm", &conv_gemm, "grouped implicit-GEMM conv / linear layer on tcgen05 (fprop, dgrad, wgrad + SGD)");
    m.def("conv_tma", &conv_tma, "TMA-fed grouped implicit-GEMM conv / linear layer on tcgen05");
    m.def("tma_encode", &tma_encode, "encode a tiled fp32 tensor map (rank <= 5)");
    m.def("set_pdl", &set_pdl, "enable / disable programmatic dependent launch for the fused tapes");
    m.def("get_pdl", &get_pdl);
    m.def("gather_grouped", &gather_grouped);
    m.def("im2col_pack", &im2col_pack, "mini-batch gather + im2col of the first layer + per-step packing of its weights");
    m.def("bn_fwd_grouped", &bn_fwd_grouped);
    m.def("bn_bwd_grouped", &bn_bwd_grouped);
    m.def("maxpool_fwd_grouped", &maxpool_fwd_grouped);
    m.def("maxpool_bwd_grouped", &maxpool_bwd_grouped);
    m.def("avgpool_grouped", &avgpool_grouped);
    m.def("dropout_grouped", &dropout_grouped);
    m.def("loss_grouped", &loss_grouped);
}
