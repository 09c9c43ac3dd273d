// TEST INFRASTRUCTURE: mutation fuzz of the library's bitstream gather (vvc_intra_b200/csrc/vvcb_gather.inc), built with
// -fsanitize=address,undefined by tests/test_assemble.py.  The C ABI takes files from outside: a damaged stream must come
// back as a status code, never as a fault.     usage: gather_fuzz <iterations> <workdir> <stream> [<stream> ...]
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <random>
#include "../../vvc_intra_b200/csrc/vvcb_gather.inc"

static std::vector<uint8_t> slurp(const char* p)
{
  std::vector<uint8_t> b;
  FILE* f = fopen(p, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", p); exit(2); }
  int c;
  while ((c = fgetc(f)) != EOF) b.push_back((uint8_t)c);
  fclose(f);
  return b;
}

int main(int argc, char** argv)
{
  if (argc < 4) return 2;
  const int iterations = atoi(argv[1]);
  const std::string in = std::string(argv[2]) + "/fuzz_in.bin", out = std::string(argv[2]) + "/fuzz_out.bin";
  std::mt19937 rng(1234);
  int ok = 0, refused = 0, unsupported = 0;
  for (int it = 0; it < iterations; it++) {
    std::vector<uint8_t> b = slurp(argv[3 + rng() % (argc - 3)]);
    const int kind = rng() % 4, edits = 1 + rng() % 4;
    for (int m = 0; m < edits && !b.empty(); m++) {
      const size_t at = rng() % b.size();
      if (kind == 0) b[at] ^= (uint8_t)(1u << (rng() % 8));            // bit flip
      else if (kind == 1) b[at] = (uint8_t)rng();                      // byte replaced
      else if (kind == 2) b.resize(at);                                // truncated
      else b.insert(b.begin() + at, (uint8_t)(rng() % 4));             // small byte inserted (start codes, emulation prevention)
    }
    FILE* f = fopen(in.c_str(), "wb");
    if (!f) return 2;
    fwrite(b.data(), 1, b.size(), f);
    fclose(f);
    const char* paths[2] = {in.c_str(), in.c_str()};
    char err[128];
    int k = 0;
    const int rc[2] = {vvcb_gather_sequential(paths, 2, out.c_str(), 1, nullptr, err, sizeof err), vvcb_gather_parcat(paths, 2, out.c_str(), &k, err, sizeof err)};
    for (int r : rc) {
      if (r == VVCB_OK) ok++; else if (r == VVCB_ERR_ARG) refused++; else if (r == VVCB_ERR_STATE) unsupported++; else return 3;
    }
  }
  printf("%d %d %d\n", ok, refused, unsupported);
  return 0;
}
