// raven-b200: drop-in `bioparser/fasta_parser.hpp` (see bioparser/parser.hpp).
// FASTA: '>' header line, sequence over any number of lines.
#ifndef BIOPARSER_FASTA_PARSER_HPP_
#define BIOPARSER_FASTA_PARSER_HPP_

#include "bioparser/parser.hpp"

namespace bioparser {

template <class T>
class FastaParser : public Parser<T> {
 public:
  std::vector<std::unique_ptr<T>> Parse(std::uint64_t bytes,
                                        bool shorten_names = true) override {
    std::vector<std::unique_ptr<T>> dst;
    std::uint64_t parsed = 0;
    std::string line, header, data;
    bool in_record = false;
    auto flush = [&]() {
      if (!in_record) return;
      const std::uint32_t name_len = this->NameLength(header, shorten_names);
      if (name_len == 0 || data.empty()) {
        throw std::invalid_argument(
            "[bioparser::FastaParser] error: invalid file format");
      }
      dst.emplace_back(std::unique_ptr<T>(new T(header.c_str() + 1, name_len, data.c_str(),
                                                static_cast<std::uint32_t>(data.size()))));
      parsed += header.size() + data.size();
      in_record = false;
    };
    while (true) {
      bool got;
      if (this->has_pending_) {
        line.swap(this->pending_);
        this->has_pending_ = false;
        got = true;
      } else {
        got = this->ReadLine(&line);
      }
      if (!got) break;
      if (!line.empty() && line[0] == '>') {
        flush();
        if (parsed >= bytes) {  // enough for this call: keep the header for the next
          this->pending_ = line;
          this->has_pending_ = true;
          return dst;
        }
        header = line;
        data.clear();
        in_record = true;
      } else {
        if (!in_record) {
          if (line.empty()) continue;
          throw std::invalid_argument(
              "[bioparser::FastaParser] error: invalid file format");
        }
        this->StripSpaces(&line);
        data += line;
      }
    }
    flush();
    return dst;
  }

 private:
  friend Parser<T>;
  explicit FastaParser(gzFile file) : Parser<T>(file) {}
};

}  // namespace bioparser

#endif  // BIOPARSER_FASTA_PARSER_HPP_
