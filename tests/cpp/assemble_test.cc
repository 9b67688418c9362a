// RavenTest.Assemble (RavenTest/src/raven_test.cpp:50-67) with the reference's
// own ConstructGraph / Assemble / Polish / GetUnitigs sources (compiled in
// place) running on the B200 engine through our drop-in facades
// (ram::MinimizerEngine, racon::Polisher, edlib). Writes the unitigs for
// tests/test_gpu_dropin.py, which compares them with the CPU oracle pipeline.
//   usage: assemble_test <reads.bin> <out.txt> <minhash> <rounds>
#include <atomic>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "raven/graph/assemble.h"
#include "raven/graph/common.h"
#include "raven/graph/construct.h"
#include "raven/graph/polish.hpp"
#include "raven/graph/serialization/binary.h"

std::atomic<std::uint32_t> biosoup::NucleicAcid::num_objects{0};

namespace raven {
void StoreGraphToFile(const Graph&) { throw std::logic_error("no checkpoints here"); }
}  // namespace raven

template <typename T>
static std::vector<T> ReadVec(std::ifstream& f) {
  std::uint64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), 8);
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  auto words = ReadVec<std::uint64_t>(in);
  auto woff = ReadVec<std::uint64_t>(in);
  auto lens = ReadVec<std::uint32_t>(in);
  auto bq = ReadVec<std::uint8_t>(in);
  auto bqoff = ReadVec<std::uint64_t>(in);
  std::vector<std::unique_ptr<biosoup::NucleicAcid>> seqs;
  for (std::size_t i = 0; i < lens.size(); ++i) {
    auto s = std::make_unique<biosoup::NucleicAcid>();
    s->id = i;
    s->name = std::to_string(i);
    s->deflated_data.assign(words.begin() + woff[i], words.begin() + woff[i + 1]);
    s->inflated_len = lens[i];
    s->is_reverse_complement = false;
    if (!bqoff.empty()) s->block_quality.assign(bq.begin() + bqoff[i], bq.begin() + bqoff[i + 1]);
    seqs.emplace_back(std::move(s));
  }
  biosoup::NucleicAcid::num_objects = seqs.size();
  try {
    raven::Graph graph;
    auto pool = std::make_shared<thread_pool::ThreadPool>(8);
    raven::OverlapPhaseCfg cfg{};
    cfg.useMinhash = std::stoi(argv[3]) != 0;
    raven::ConstructGraph(graph, seqs, pool, false, cfg);
    raven::Assemble(pool, graph, false);
    raven::PolishCfg pcfg{};
    pcfg.num_rounds = std::stoul(argv[4]);
    raven::Polish(pool, graph, false, seqs, pcfg);
    std::ofstream out(argv[2]);
    for (const auto& u : raven::GetUnitigs(graph)) {
      out << u->name << "\n" << u->InflateData() << "\n";
    }
  } catch (const std::exception& e) {
    std::cerr << "assemble_test: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
