#!/usr/bin/env python
"""profiles/r02_sanitizer_summary.txt from the three compute-sanitizer logs of tools/sanitize_targets.py (gpurun_out/r2_san_*_final.txt)"""
import re, subprocess, sys
def ok_line(path):
    return "\n".join(l[:400] for l in open(path, errors="replace").read().splitlines() if "SANITIZE_TARGETS_OK" in l or "SUMMARY" in l)
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
sites = {}
for l in open("gpurun_out/r2_san_racecheck_final.txt", errors="replace"):
    if "Error: Race reported" in l:
        k = re.sub(r"\+0x[0-9a-f]+", "", l.strip())[:230]
        sites[k] = sites.get(k, 0) + 1
print(f"""# compute-sanitizer on tools/sanitize_targets.py (B200, CUDA 12.9), round 2, final code (commit {head})
# the targets: every tuned kernel incl. two-segment / unaligned / ragged cases, the launch-parameter ring forms (256 taps, 16-warp
# real decimators), the fused FM stages, zero-copy held pushes and the persistent consumer; each case also compares its result
# with the generic kernel or a second run (checksums)

## memcheck (2^21-sample ragged streams)
{ok_line("gpurun_out/r2_san_memcheck_final.txt")}

## synccheck (2^17-sample streams)
{ok_line("gpurun_out/r2_san_synccheck_final.txt")}

## racecheck (2^17-sample streams): hazards by site (--print-limit 20)""")
for k, v in sorted(sites.items(), key=lambda t: -t[1]):
    print(f"{v:7d} {k}")
print(ok_line("gpurun_out/r2_san_racecheck_final.txt"))
print("""
Reading of the racecheck report.  Two source sites only, both by design:
 * gen_publish / gen_read (ring_common.cuh): the generation guard is a polled shared-memory word (volatile store by the
   filler, volatile load by consumers) -- racecheck flags every polled flag as a write/read hazard.  The consumer never
   uses the word for anything but deciding whether to wait again; the data itself is ordered by the mbarrier.
 * the zero-fill stores of the ring decimator's edge fill (kernels_fast.cu, issue_fill_edge; one line of source, reported once
   per instantiation that ran a ragged case) against the window loads: ordered by __syncwarp + mbarrier.arrive.expect_tx
   (release) / mbarrier.try_wait (acquire); racecheck does not model inline-PTX mbarrier synchronisation (it reports nothing
   for the TMA bulk copies either: the async proxy is invisible to it).
Nothing is reported for the fused FM front end (in-kernel boundary pass, ticket counter), the low-rate kernel, the contiguous-slot
rings or the persistent consumer.

What the sanitizers did NOT find: the persistent consumer's intermittent "unspecified launch failure" (1 in ~40 passes of
32768 vector pushes; never under memcheck, whose slowdown keeps the kernel behind the host).  It was a protocol bug on the
path taken when the kernel runs AHEAD of the host -- a parity wait on a slot's `empty` barrier that aliased with the phase two
back, plus a tile left unfilled when its previous generation was still unclaimed (DESIGN.md section 4.1b) -- found by forcing
that path (SDR_B200_PERSIST_FLAGS=1: deterministic failure) and fixed (kernels_fast.cu, claim_and_fill).
tests/test_gpu_persistent.py now runs the suite in that mode too, and the soak (r02_soak.txt) feeds the consumer 200 x 32768
vector pushes: 0 mismatching passes.""")
