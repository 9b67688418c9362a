// CPU check of the drop-in I/O dependencies the reference's CLI needs:
// bioparser (include/bioparser/*.hpp on zlib; RavenLib/src/io.cc:7-41,
// RavenExe/src/main.cc:258-272) and the cereal archives (include/cereal;
// RavenLib/src/binary.cc:73-93, graph_repr.cc:400-416).
//   usage: io_test parse <fasta|fastq> <path> [chunk_bytes]   -> one line per record
//          io_test cereal <tmp path>                            -> "ok" or throws
#include <atomic>
#include <fstream>
#include <iostream>
#include <sstream>

#include "bioparser/fasta_parser.hpp"
#include "bioparser/fastq_parser.hpp"
#include "biosoup/nucleic_acid.hpp"
#include "cereal/archives/binary.hpp"
#include "cereal/archives/json.hpp"
#include "cereal/types/memory.hpp"
#include "cereal/types/vector.hpp"

std::atomic<std::uint32_t> biosoup::NucleicAcid::num_objects{0};

namespace demo {
struct Leaf {
  std::uint32_t id = 0;
  double w = 0;
  std::vector<bool> bits;
  template <class Archive>
  void serialize(Archive& ar) { ar(CEREAL_NVP(id), CEREAL_NVP(w), CEREAL_NVP(bits)); }
};
struct Free {
  std::string name;
  std::vector<std::uint16_t> data;
};
template <class Archive>
void serialize(Archive& ar, Free& f) { ar(f.name, f.data); }
struct Root {
  int stage = 0;
  std::vector<std::unique_ptr<Leaf>> leaves;
  std::vector<std::pair<std::uint32_t, std::uint32_t>> regions;
  std::unordered_set<std::uint32_t> set;
  Free free_member;
  bool flag = false;
  template <class Archive>
  void serialize(Archive& ar) { ar(stage, leaves, regions, set, free_member, flag); }
};
}  // namespace demo

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "";
  if (mode == "parse" && argc >= 4) {
    using P = bioparser::Parser<biosoup::NucleicAcid>;
    std::unique_ptr<P> p;
    try {
      p = std::string(argv[2]) == "fasta" ? P::Create<bioparser::FastaParser>(argv[3])
                                           : P::Create<bioparser::FastqParser>(argv[3]);
      const std::uint64_t chunk = argc > 4 ? std::stoull(argv[4]) : static_cast<std::uint64_t>(-1);
      while (true) {
        auto v = p->Parse(chunk);
        if (v.empty()) break;
        for (const auto& s : v) {
          std::cout << s->id << "\t" << s->name << "\t" << s->InflateData() << "\t";
          for (auto q : s->block_quality) std::cout << static_cast<int>(q) << ",";
          std::cout << "\n";
        }
      }
    } catch (const std::invalid_argument& e) {
      std::cout << "invalid_argument: " << e.what() << "\n";
      return 3;
    }
    return 0;
  }
  if (mode == "cereal" && argc >= 3) {
    demo::Root a;
    a.stage = -3;
    for (int i = 0; i < 5; ++i) {
      if (i == 2) {
        a.leaves.emplace_back();  // a null pointer survives
        continue;
      }
      a.leaves.emplace_back(new demo::Leaf());
      a.leaves.back()->id = 10 + i;
      a.leaves.back()->w = 0.5 * i;
      a.leaves.back()->bits.assign(70 + i, false);
      a.leaves.back()->bits[i] = a.leaves.back()->bits[69] = true;
    }
    a.regions = {{1, 2}, {30, 40}};
    a.set = {7, 9, 11};
    a.free_member.name = "free \"quoted\"";
    a.free_member.data = {1, 2, 65535};
    a.flag = true;
    {
      std::ofstream os(argv[2], std::ios::binary);
      cereal::BinaryOutputArchive ar(os);
      ar(a);
    }
    demo::Root b;
    {
      std::ifstream is(argv[2], std::ios::binary);
      cereal::BinaryInputArchive ar(is);
      ar(b);
    }
    bool ok = b.stage == -3 && b.leaves.size() == 5 && !b.leaves[2] && b.leaves[4]->id == 14 &&
              b.leaves[3]->w == 1.5 && b.leaves[1]->bits.size() == 71 && b.leaves[1]->bits[1] &&
              b.leaves[1]->bits[69] && !b.leaves[1]->bits[2] && b.regions == a.regions &&
              b.set == a.set && b.free_member.name == a.free_member.name &&
              b.free_member.data == a.free_member.data && b.flag;
    {  // a truncated archive must throw
      std::stringstream half;
      {
        cereal::BinaryOutputArchive ar(half);
        ar(a);
      }
      std::string bytes = half.str();
      std::stringstream cut(bytes.substr(0, bytes.size() / 2));
      cereal::BinaryInputArchive ar(cut);
      demo::Root c;
      bool threw = false;
      try {
        ar(c);
      } catch (const std::exception&) {
        threw = true;
      }
      ok = ok && threw;
    }
    std::ostringstream js;
    {
      cereal::JSONOutputArchive ar(js);
      ar(cereal::make_nvp(std::to_string(42), *a.leaves[0]));
      ar(cereal::make_nvp("root", a));
    }
    std::cout << js.str();
    std::cout << (ok ? "ok" : "MISMATCH") << "\n";
    return ok ? 0 : 1;
  }
  std::cerr << "usage: io_test parse fasta|fastq path [chunk] | io_test cereal tmp\n";
  return 2;
}
