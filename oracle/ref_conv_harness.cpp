// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
// The reference's own text -> CSR converter (src/converters/converter.cc: Converter(file_type, file_name, is_bipartite) and
// generate_binary_graph -> GraphT::write_to_file, src/common/graph.cc:467-508) behind a main(): the reference's main.cc only serves
// the "gr" split path (its constructor call is commented out, main.cc:19-20). build_ref.sh compiles this file with the reference
// sources where they lie into oracle/_ref/ref_convert.
//   ref_convert <mtx|edges|lg> <input file> <output prefix> [is_bipartite]
#include "converter.h"

int main(int argc, char* argv[]) {
  if (argc < 4) { printf("usage: %s <mtx|edges|lg> <input> <out_prefix> [is_bipartite]\n", argv[0]); return 1; }
  Converter converter(argv[1], argv[2], argc > 4 && atoi(argv[4]) != 0);
  converter.generate_binary_graph(argv[3], true, true, false, false);
  return 0;
}
