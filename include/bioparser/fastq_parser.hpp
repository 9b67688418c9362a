// raven-b200: drop-in `bioparser/fastq_parser.hpp` (see bioparser/parser.hpp).
// FASTQ: '@' header, sequence lines, '+' line, as many quality characters as
// bases (multi-line records are accepted).
#ifndef BIOPARSER_FASTQ_PARSER_HPP_
#define BIOPARSER_FASTQ_PARSER_HPP_

#include "bioparser/parser.hpp"

namespace bioparser {

template <class T>
class FastqParser : public Parser<T> {
 public:
  std::vector<std::unique_ptr<T>> Parse(std::uint64_t bytes,
                                        bool shorten_names = true) override {
    std::vector<std::unique_ptr<T>> dst;
    std::uint64_t parsed = 0;
    std::string header, line, data, quality;
    auto invalid = []() {
      return std::invalid_argument("[bioparser::FastqParser] error: invalid file format");
    };
    while (parsed < bytes) {
      // header
      bool got = this->ReadLine(&header);
      while (got && header.empty()) got = this->ReadLine(&header);
      if (!got) break;
      if (header[0] != '@') throw invalid();
      // sequence up to the '+' line
      data.clear();
      while (true) {
        if (!this->ReadLine(&line)) throw invalid();
        if (!line.empty() && line[0] == '+') break;
        this->StripSpaces(&line);
        data += line;
      }
      // quality: as many characters as bases
      quality.clear();
      while (quality.size() < data.size()) {
        if (!this->ReadLine(&line)) throw invalid();
        this->StripSpaces(&line);
        quality += line;
      }
      const std::uint32_t name_len = this->NameLength(header, shorten_names);
      if (name_len == 0 || data.empty() || quality.size() != data.size()) throw invalid();
      dst.emplace_back(std::unique_ptr<T>(
          new T(header.c_str() + 1, name_len, data.c_str(), static_cast<std::uint32_t>(data.size()),
                quality.c_str(), static_cast<std::uint32_t>(quality.size()))));
      parsed += header.size() + 2 * data.size();
    }
    return dst;
  }

 private:
  friend Parser<T>;
  explicit FastqParser(gzFile file) : Parser<T>(file) {}
};

}  // namespace bioparser

#endif  // BIOPARSER_FASTQ_PARSER_HPP_
