// gpu_converter: text graph -> the reference's binary CSR files, CSR construction on the device.
//   ./gpu_converter <mtx|edges|lg> <input_file> <output_prefix> [is_bipartite] [write_meta]
// (the reference's own main, src/converters/main.cc, only serves the "gr" split path; its constructor call is commented out)
#include <cstdio>
#include <cstdlib>
#include "gai_converter.h"

int main(int argc, char* argv[]) {
  if (argc < 4) {
    printf("Usage: %s <mtx|edges|lg> <input_file> <output_prefix> [is_bipartite(0)] [write_meta(1)]\n", argv[0]);
    return 1;
  }
  Converter converter(argv[1], argv[2], argc > 4 && atoi(argv[4]) != 0);
  converter.generate_binary_graph(argv[3], true, true, false, false);
  if (argc <= 5 || atoi(argv[5]) != 0) converter.write_meta(argv[3]);
  return 0;
}
