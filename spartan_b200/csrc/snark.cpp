// spartan_b200 — SNARK path (placeholder until the SPARK driver lands; every entry point fails loudly).
#include "snark.hpp"
#include "../../include/spartan_b200.h"

namespace sp {
SnarkGens::SnarkGens(Ctx*, size_t, size_t, size_t, size_t) { throw SpError(SP_ERR_INTERNAL, "SNARK path not built yet"); }
void SnarkEncoding::ser_commitment(Writer&) const {}
void snark_encode(Ctx&, const Instance&, const SnarkGens&, SnarkEncoding&) { throw SpError(SP_ERR_INTERNAL, "SNARK path not built yet"); }
void snark_prove(Ctx&, const Instance&, const SnarkEncoding&, const u256*, const std::vector<Fq>&, const SnarkGens&, Transcript&, const Fq&, Writer&) {
  throw SpError(SP_ERR_INTERNAL, "SNARK path not built yet");
}
}  // namespace sp
