# multi-GPU experiment driver: panel exchange (NCCL broadcast vs P2P pull), T^B chunking / SM reservation
G=${G:-2}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $G --steps 2 --warmup 1 --no-cpu > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python -c "
import json;d=json.load(open('gpurun_out/x_$name.json'));p=d['phases_ms'];print('$name', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'factor', round(d['host_call_ms']['factor'],1), 'trailing', round(p['trailing_ms'],1), 'comm', round(p['comm_ms'],1), 'chain', round(p['panel_chain_ms'],1), 'predict', round(p['predict_ms'],1), 'parity', d.get('parity_vs_n1',{}).get('ok'))" || tail -5 gpurun_out/x_$name.err
}
for v in "$@"; do
  case $v in
    p2p) run p2p ;;
    nccl) run nccl SB_P2P=0 ;;
    chunk*) run p2p_$v SB_OZ_CHUNK=${v#chunk} ;;
    sms*) run p2p_$v SB_LOOKAHEAD_SMS=${v#sms} ;;
  esac
done
