// dais_replay.cuh -- replay of DAIS adder-graph programs on the device.
//
// Bit-exact port of the parts of the reference interpreter that the CMVM path can produce
// (reference src/da4ml/_binary/dais/DAISInterpreter.cc): opcode -1 (input load + wrap, :300-308, quantize :139-152),
// opcodes 0/1 (shift_add :114-137), output scaling (inference :394-405).  Programs are the int32 words written by
// CombLogic.to_binary (reference types.py:500-541).  Ops are levelised on the host (an op's level = 1 + max level of
// its operands); one launch per level, thread = (op of the level, sample) with the sample index fastest, so all
// buffer traffic is coalesced: this kernel family is plain HBM-bound streaming (24 B per op per sample).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace da {

struct DaisOp { // layout of the reference's dais::Op (DAISInterpreter.hh:44-51)
    int32_t opcode, id0, id1, data_low, data_high;
    int32_t is_signed, integers, fractionals;
};

__device__ __forceinline__ int dais_width(const DaisOp &o) { return o.integers + o.fractionals + (o.is_signed ? 1 : 0); }

// DAISInterpreter::quantize with dtype_from == dtype_to (wrap into the op's range)
__device__ __forceinline__ long long dais_wrap(long long value, const DaisOp &o) {
    const int w = dais_width(o);
    const int int_min = o.is_signed ? -(1 << (w - 1)) : 0; // int32 arithmetic, as the reference
    const long long mod = 1LL << w;
    const long long av = value < 0 ? -value : value;
    return ((value - int_min + (av / mod + 1) * mod) % mod) + int_min;
}

// buffer layout: [n_ops][S] int64, S = samples of the current chunk
__global__ void __launch_bounds__(256) dais_level_kernel(const DaisOp *ops, const int *order, int n_level_ops, const int *inp_shifts, const double *inputs, int n_in, long long S, long long *buffer) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)n_level_ops * S)
        return;
    const int k = (int)(tid / S);
    const long long s = tid - (long long)k * S;
    const int i = order[k];
    const DaisOp op = ops[i];
    long long r;
    if (op.opcode == -1) {
        const double x = inputs[s * n_in + op.id0];
        const long long v = (long long)floor(x * exp2((double)(inp_shifts[op.id0] + op.fractionals)));
        r = dais_wrap(v, op);
    }
    else {
        const DaisOp o0 = ops[op.id0], o1 = ops[op.id1];
        const long long v1 = buffer[(long long)op.id0 * S + s];
        long long v2 = buffer[(long long)op.id1 * S + s];
        if (op.opcode == 1)
            v2 = -v2;
        const int shift = op.data_low;
        const int actual = shift + o0.fractionals - o1.fractionals;
        // shifts of negative values: two's complement, as the reference's compiled code behaves
        if (actual > 0)
            r = v1 + (long long)((unsigned long long)v2 << actual);
        else
            r = (long long)((unsigned long long)v1 << -actual) + v2;
        const int gshift = max(o0.fractionals, o1.fractionals - shift) - op.fractionals;
        if (gshift > 0)
            r >>= gshift;
    }
    buffer[(long long)i * S + s] = r;
}

__global__ void __launch_bounds__(256) dais_output_kernel(const DaisOp *ops, const int *out_idxs, const int *out_shifts, const int *out_negs, int n_out, long long S, const long long *buffer, double *outputs) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)n_out * S)
        return;
    const int o = (int)(tid / S);
    const long long s = tid - (long long)o * S;
    const int idx = out_idxs[o];
    double y = 0.0;
    if (idx >= 0) {
        long long v = buffer[(long long)idx * S + s];
        if (out_negs[o])
            v = -v;
        y = (double)v * exp2((double)(out_shifts[o] - ops[idx].fractionals));
    }
    outputs[s * n_out + o] = y;
}

} // namespace da
