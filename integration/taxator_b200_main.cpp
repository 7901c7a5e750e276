// Parity harness for integration/taxonpredictionmodelgpu.hh: the reference's own `taxator` program (core/taxator.cpp,
// unmodified, compiled where it lies) with its RPA model swapped for the B200 binding.  What a maintainer would do
// by editing one identifier at core/taxator.cpp:252 is done here with the preprocessor so that the reference tree
// stays untouched: the reference's model header is included first (its include guard then keeps taxator.cpp from
// re-including it), the binding next, and from there on the name RPAPredictionModel means RPAPredictionModelB200.
// Everything else -- option parsing, taxonomy / mapping loading, the record-set generator, the -p N consumer
// threads, GFF3 printing -- is the reference's code.
#include "src/taxonpredictionmodelsequence.hh"
#include "taxonpredictionmodelgpu.hh"
#define RPAPredictionModel RPAPredictionModelB200
#include "taxator.cpp"
