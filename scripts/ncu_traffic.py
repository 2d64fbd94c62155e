"""profiles/rNN_step_ncu_full_summary.json (scripts/ncu_summarize.py of an `ncu --set full` capture of
scripts/one_step.py) -> profiles/rNN_kernel_traffic.json: DRAM bytes and pipe utilisation per launch site of ONE learner
step, keyed like bench.py's per_kernel_ms.  The last complete run of len(SITES) launches starting at the pack kernel is one whole step.

Usage: python scripts/ncu_traffic.py profiles/r02_step_ncu_full_summary.json profiles/r02_kernel_traffic.json [samples] [ppo|rnd]"""
import json
import sys

# launch order of one bf16 learner step (umma_net.cu / net.cu) and a substring each kernel name must contain
SITES = [('pack_weights', 'pack_all_kernel'), ('conv1_fwd', 'conv1_fwd_i8_kernel'), ('conv2_fwd', 'tma_conv2_kernel<64'),
         ('conv3_fwd', 'tma_conv_kernel<64'), ('fc_fwd', 'tma_gemm2_kernel'), ('heads_loss', 'heads_mma_kernel'),
         ('fc_wgrad', 'wgrad_tma_kernel'), ('fc_wgrad_t', 'wgrad_untranspose_kernel'), ('fc_dgrad', 'tma_gemm_kernel<256'),
         ('conv3_wgrad', 'wgrad_conv_tma_kernel<64, 5'), ('conv3_wgrad_t', 'wgrad_untranspose_kernel'),
         ('conv3_dgrad', 'tma_conv2s_kernel<64'), ('conv2_wgrad', 'wgrad_conv_tma_kernel<64, 4'),
         ('conv2_wgrad_t', 'wgrad_untranspose_kernel'), ('conv2_dgrad', 'tma_conv_kernel<128'),
         ('conv1_wgrad', 'conv1_wgrad_s2d_kernel'), ('adam', 'adam_kernel')]
# one RND predictor step (umma_rnd.cu: rnd_bf16_forward / mse / backward, then Adam + repack)
RND_SITES = [('rnd_norm_table', 'rnd_norm_table_kernel'), ('rnd_conv1_fwd', 'rnd_conv1_fwd_kernel'), ('rnd_conv2_fwd_t', 'tma_conv_kernel<64'),
             ('rnd_conv3_fwd_t', 'tma_conv_kernel<64'), ('rnd_conv2_fwd_s', 'tma_conv_kernel<64'), ('rnd_conv3_fwd_s', 'tma_conv_kernel<64'),
             ('rnd_fc_t', 'tma_gemm'), ('rnd_fc1', 'tma_gemm'), ('rnd_fc2', 'tma_gemm'),
             ('rnd_fc3', 'tma_gemm'), ('rnd_mse', 'rnd_mse_bf16_kernel'), ('rnd_fc3_wgrad', 'wgrad_tma_kernel'),
             ('rnd_fc3_dgrad', 'tma_gemm'), ('rnd_fc2_bgrad', 'rnd_colsum_bf16_kernel'), ('rnd_fc2_wgrad', 'wgrad_tma_kernel'),
             ('rnd_fc2_dgrad', 'tma_gemm'), ('rnd_fc1_bgrad', 'rnd_colsum_bf16_kernel'), ('rnd_fc1_wgrad', 'wgrad_tma_kernel'),
             ('rnd_fc1_wgrad_t', 'wgrad_untranspose_kernel'), ('rnd_fc1_dgrad', 'tma_gemm'),
             ('rnd_conv3_wgrad', 'wgrad_conv_tma_kernel<64, 8'), ('rnd_conv3_wgrad_t', 'wgrad_untranspose_kernel'),
             ('rnd_conv3_dgrad', 'tma_conv_kernel<64'), ('rnd_conv2_wgrad', 'wgrad_conv_tma_kernel<64, 4'),
             ('rnd_conv2_wgrad_t', 'wgrad_untranspose_kernel'), ('rnd_conv2_dgrad', 'tma_conv_kernel<128'),
             ('rnd_conv1_wgrad', 'rnd_conv1_wgrad_kernel'), ('rnd_adam', 'adam_kernel'), ('rnd_pack', 'rnd_pack_kernel')]
P = 'avg.pct_of_peak_sustained_'


def main():
    global SITES
    rows = json.load(open(sys.argv[1]))
    samples = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
    if len(sys.argv) > 4 and sys.argv[4] == 'rnd':
        SITES = RND_SITES
    rows = [r for r in rows if 'memset' not in r['kernel'].lower()]
    starts = [i for i, r in enumerate(rows) if SITES[0][1] in r['kernel'] and i + len(SITES) <= len(rows)]
    rows = rows[starts[-1]:starts[-1] + len(SITES)]          # the last COMPLETE step of the capture
    out = {'samples_per_launch': samples,
           'source': 'ncu --set full --clock-control none, scripts/one_step.py (last captured step), B200', 'kernels': {}}
    for (site, pat), r in zip(SITES, rows):
        assert pat in r['kernel'], (site, pat, r['kernel'])
        rd, wr = r.get('dram__bytes_read.sum', 0.0), r.get('dram__bytes_write.sum', 0.0)
        out['kernels'][site] = {
            'kernel': r['kernel'], 'dram_read': int(rd), 'dram_write': int(wr), 'dram_bytes': int(rd + wr),
            'ncu_us': round(r.get('gpu__time_duration.sum', 0.0), 1),
            'tensor_pipe_pct': round(r.get('sm__pipe_tensor_cycles_active.' + P + 'active', 0.0), 1),
            'l1tex_pct': round(r.get('l1tex__throughput.' + P + 'elapsed', 0.0), 1),
            'lts_pct': round(r.get('lts__throughput.' + P + 'elapsed', 0.0), 1),
            'sm_pct': round(r.get('sm__throughput.' + P + 'elapsed', 0.0), 1),
            'regs': int(r.get('launch__registers_per_thread', 0)), 'grid': int(r.get('launch__grid_size', 0)),
            'block': int(r.get('launch__block_size', 0))}
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
    for k, v in out['kernels'].items():
        print(f"{k:14s} {v['ncu_us']:8.1f} us  dram {v['dram_bytes'] / 1e6:9.1f} MB  tensor {v['tensor_pipe_pct']:5.1f}%  l1tex {v['l1tex_pct']:5.1f}%")


if __name__ == '__main__':
    main()
