// oracle/dais_glue.cc -- TEST INFRASTRUCTURE ONLY.  ctypes window onto the reference's DAIS interpreter
// (/root/reference/src/da4ml/_binary/dais/DAISInterpreter.{hh,cc}, compiled in place, unmodified); the loop below is
// `_run_interp` of the reference's dais/bindings.cc:14-30 without the nanobind types.
#include "DAISInterpreter.hh"
#include <string>

static thread_local std::string g_err;

extern "C" {
const char *ref_dais_last_error() { return g_err.c_str(); }
int ref_dais_run(const int32_t *program, int64_t n_words, const double *inputs, int64_t n_samples, double *outputs) {
    try {
        const std::span<const int32_t> bin(program, (size_t)n_words);
        const int32_t n_in = program[2], n_out = program[3];
        dais::DAISInterpreter interp;
        interp.load_from_binary(bin);
        for (int64_t i = 0; i < n_samples; ++i) {
            const std::span<const double> inp(&inputs[i * n_in], (size_t)n_in);
            std::span<double> out(&outputs[i * n_out], (size_t)n_out);
            interp.inference(inp, out);
        }
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
}
