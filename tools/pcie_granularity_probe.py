import time, torch
dev = torch.device('cuda:0')
def run(nbytes_each, n_in, n_out, chunks, streams=3):
    hin = [torch.empty(nbytes_each*chunks, dtype=torch.uint8, pin_memory=True) for _ in range(n_in)]
    hout = [torch.empty(nbytes_each*chunks, dtype=torch.uint8, pin_memory=True) for _ in range(n_out)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    din = [[torch.empty(nbytes_each, dtype=torch.uint8, device=dev) for _ in range(n_in)] for _ in range(streams)]
    dout = [[torch.empty(nbytes_each, dtype=torch.uint8, device=dev) for _ in range(n_out)] for _ in range(streams)]
    def go():
        for c in range(chunks):
            s = c % streams
            with torch.cuda.stream(ss[s]):
                for k in range(n_in):
                    din[s][k].copy_(hin[k][c*nbytes_each:(c+1)*nbytes_each], non_blocking=True)
                for k in range(n_out):
                    hout[k][c*nbytes_each:(c+1)*nbytes_each].copy_(dout[s][k], non_blocking=True)
        torch.cuda.synchronize()
    go()
    t0 = time.perf_counter(); go(); t = time.perf_counter() - t0
    up = nbytes_each*n_in*chunks/t/1e9; dn = nbytes_each*n_out*chunks/t/1e9
    print(f"{nbytes_each>>20} MiB copies, {n_in} in / {n_out} out per chunk, {streams} streams: up {up:.1f} GB/s, down {dn:.1f} GB/s")
run(4<<20, 16, 12, 32)
run(4<<20, 16, 12, 32, streams=2)
run(8<<20, 16, 12, 16)
run(4<<20, 16, 0, 32)
run(4<<20, 16, 16, 32)
run(64<<20, 1, 1, 32)
