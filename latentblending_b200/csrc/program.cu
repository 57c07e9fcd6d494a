// program.cu -- the UNet graph executor: a flat, pre-validated list of kernel launches.
//
// The host side (latentblending_b200/unet.py) lowers one SDXL-UNet forward for a fixed
// (batch, height, width) into ~1.7k lb_op records over static device buffers; this file
// turns them into prepared launches once (TMA descriptors encoded, tilings chosen) and
// replays them on a stream with no Python, no allocation and no host sync in the loop.
// Replaces the eager PyTorch module walk behind pipe.unet(...) (diffusers_holder.py:336-344).
#include <vector>

#include "gemm_sm100.cuh"

struct AttnPlan;
int attn_plan_build_opaque(lb_ctx* ctx, const lb_attn_desc& d, void** plan_out);
int attn_plan_launch_opaque(void* plan, cudaStream_t st);
void attn_plan_free_opaque(void* plan);

struct lb_program {
    lb_ctx* ctx;
    struct Node {
        lb_op op;
        GemmPlan gemm;
        void* attn;
    };
    std::vector<Node> nodes;
};

extern "C" int lb_program_create(lb_ctx* ctx, const lb_op* ops, int64_t n_ops, lb_program** out) {
    LB_REQUIRE(ctx && ops && out && n_ops >= 0, "lb_program_create: bad arguments");
    lb_program* prog = new lb_program();
    prog->ctx = ctx;
    prog->nodes.resize((size_t)n_ops);
    for (int64_t i = 0; i < n_ops; ++i) {
        lb_program::Node& nd = prog->nodes[(size_t)i];
        nd.op = ops[i];
        nd.attn = nullptr;
        int e = 0;
        switch (ops[i].kind) {
            case LB_OP_GEMM:
                e = gemm_plan_build(ctx, *reinterpret_cast<const GemmDesc*>(&ops[i].u.gemm), &nd.gemm);
                break;
            case LB_OP_ATTENTION:
                e = attn_plan_build_opaque(ctx, ops[i].u.attn, &nd.attn);
                break;
            case LB_OP_EMBED_INPUTS: case LB_OP_LINEAR_SMALL: case LB_OP_CONV_IN: case LB_OP_CONV_OUT:
            case LB_OP_UPSAMPLE2X: case LB_OP_IM2COL_S2: case LB_OP_GROUPNORM: case LB_OP_LAYERNORM:
            case LB_OP_LATENT_PREP: case LB_OP_SOFTMAX_ROWS: case LB_OP_POSTPROCESS_U8:
            case LB_OP_LPIPS_IM2COL_U8: case LB_OP_IM2COL: case LB_OP_MAXPOOL3S2:
                break;
            default:
                lb_set_error("lb_program_create: op %lld has unknown kind %d", (long long)i, ops[i].kind);
                e = 2;
        }
        if (e) {
            char msg[1024];
            snprintf(msg, sizeof(msg), "op %lld (kind %d): %s", (long long)i, ops[i].kind, lb_last_error());
            lb_set_error("%s", msg);
            for (auto& n2 : prog->nodes) if (n2.attn) attn_plan_free_opaque(n2.attn);
            delete prog;
            return e;
        }
    }
    *out = prog;
    return 0;
}

extern "C" int lb_program_destroy(lb_program* prog) {
    if (prog) {
        for (auto& nd : prog->nodes) if (nd.attn) attn_plan_free_opaque(nd.attn);
        delete prog;
    }
    return 0;
}

extern "C" int64_t lb_program_num_launches(lb_program* prog) {
    if (!prog) return -1;
    int64_t n = 0;
    for (auto& nd : prog->nodes) n += (nd.op.kind == LB_OP_GROUPNORM) ? 2 : 1;
    return n;
}

extern "C" int lb_program_run(lb_program* prog, float t, void* stream) {
    return lb_program_run_kinds(prog, t, 0xFFFFFFFFu, stream);
}

extern "C" int64_t lb_program_count_kinds(lb_program* prog, uint32_t kind_mask) {
    if (!prog) return -1;
    int64_t n = 0;
    for (auto& nd : prog->nodes)
        if (kind_mask & (1u << nd.op.kind)) n += (nd.op.kind == LB_OP_GROUPNORM) ? 2 : 1;
    return n;
}

extern "C" int lb_program_run_kinds(lb_program* prog, float t, uint32_t kind_mask, void* stream) {
    LB_REQUIRE(prog != nullptr, "lb_program_run: null program");
    lb_ctx* ctx = prog->ctx;
    cudaStream_t st = lb_stream(stream);
    for (size_t i = 0; i < prog->nodes.size(); ++i) {
        lb_program::Node& nd = prog->nodes[i];
        const lb_op& o = nd.op;
        if (!(kind_mask & (1u << o.kind))) continue;
        int e = 0;
        switch (o.kind) {
            case LB_OP_GEMM: e = gemm_plan_launch(nd.gemm, st); break;
            case LB_OP_ATTENTION: e = attn_plan_launch_opaque(nd.attn, st); break;
            case LB_OP_EMBED_INPUTS: {
                const auto& a = o.u.embed;
                e = lb_embed_inputs(ctx, t, a.text_embeds, a.time_ids, a.B, a.dim_t, a.pooled, a.dim_a, a.temb_in,
                                    a.add_in, stream);
                break;
            }
            case LB_OP_LINEAR_SMALL: {
                const auto& a = o.u.lin;
                e = lb_linear_small(ctx, a.x, a.ldx, a.M, a.K, a.w, a.ldw, a.bias, a.addend, a.ldadd, a.act_in,
                                    a.act_out, a.out, a.ldo, a.N, stream);
                break;
            }
            case LB_OP_CONV_IN: {
                const auto& a = o.u.conv;
                e = lb_conv_in(ctx, a.x, a.B, a.Cin, a.H, a.W, a.w, a.bias, a.Cout, a.out, a.ld_out, stream);
                break;
            }
            case LB_OP_CONV_OUT: {
                const auto& a = o.u.conv;
                e = lb_conv_out(ctx, a.x, a.ld_x, a.B, a.Cin, a.H, a.W, a.w, a.bias, a.Cout, a.out, stream);
                break;
            }
            case LB_OP_UPSAMPLE2X: {
                const auto& a = o.u.resample;
                e = lb_upsample2x(ctx, a.x, a.ld_x, a.B, a.H, a.W, a.C, a.out, a.ld_out, stream);
                break;
            }
            case LB_OP_IM2COL_S2: {
                const auto& a = o.u.resample;
                e = lb_im2col_s2(ctx, a.x, a.ld_x, a.B, a.H, a.W, a.C, a.out, stream);
                break;
            }
            case LB_OP_GROUPNORM: {
                const auto& a = o.u.norm;
                e = lb_groupnorm(ctx, a.x, a.ld_x, a.B, (int)a.rows, a.C, a.groups, a.gamma, a.beta, a.eps, a.silu,
                                 a.out, a.ld_out, a.workspace, stream);
                break;
            }
            case LB_OP_LAYERNORM: {
                const auto& a = o.u.norm;
                e = lb_layernorm(ctx, a.x, a.ld_x, a.rows, a.C, a.gamma, a.beta, a.eps, a.out, a.ld_out, stream);
                break;
            }
            case LB_OP_LATENT_PREP: {
                const auto& a = o.u.aux;
                e = lb_latent_prep(ctx, a.x, a.B, a.C, a.n, a.w, a.bias, a.out, stream);
                break;
            }
            case LB_OP_SOFTMAX_ROWS: {
                const auto& a = o.u.aux;
                e = lb_softmax_rows(ctx, a.x, a.ld_x, a.n, a.C, a.out, a.ld_out, stream);
                break;
            }
            case LB_OP_POSTPROCESS_U8: {
                const auto& a = o.u.aux;
                e = lb_postprocess_u8(ctx, a.x, a.B, a.C, a.n, a.out, (int*)const_cast<void*>(a.w), stream);
                break;
            }
            case LB_OP_LPIPS_IM2COL_U8: {
                const auto& a = o.u.patch;
                e = lb_lpips_im2col_u8(ctx, a.x, a.H, a.W, a.k, a.stride, a.pad, a.f, a.f + 3, a.out, a.C, stream);
                break;
            }
            case LB_OP_IM2COL: {
                const auto& a = o.u.patch;
                e = lb_im2col(ctx, a.x, a.ld_x, a.H, a.W, a.C, a.k, a.stride, a.pad, a.out, stream);
                break;
            }
            case LB_OP_MAXPOOL3S2: {
                const auto& a = o.u.patch;
                e = lb_maxpool3s2(ctx, a.x, a.ld_x, a.H, a.W, a.C, a.out, a.ld_out, stream);
                break;
            }
        }
        if (e) return e;
    }
    return 0;
}
