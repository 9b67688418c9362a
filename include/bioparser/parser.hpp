// raven-b200: drop-in `bioparser/parser.hpp` (un-vendored dependency, pinned
// 3.0.13 in Raven.deps.cmake:4-9). Surface the reference uses
// (RavenLib/src/io.cc:17-33, RavenExe/src/main.cc:258-272, RavenTest/src/
// raven_test.cpp:28-33):
//   bioparser::Parser<T>::Create<bioparser::FastaParser|FastqParser>(path)
//       -> std::unique_ptr<Parser<T>>; throws std::invalid_argument if the file
//          cannot be opened
//   parser->Parse(bytes /* -1 = everything */) -> std::vector<std::unique_ptr<T>>
//          T is built as T(name, name_len, data, data_len[, quality, quality_len]);
//          names are cut at the first white space; throws std::invalid_argument
//          on a malformed file
//   parser->Reset()
// Own implementation on zlib (gzread handles plain and gzip files alike).
#ifndef BIOPARSER_PARSER_HPP_
#define BIOPARSER_PARSER_HPP_

#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace bioparser {

template <class T>
class Parser {
 public:
  Parser(const Parser&) = delete;
  Parser& operator=(const Parser&) = delete;

  virtual ~Parser() {
    if (file_) gzclose(file_);
  }

  template <template <class> class P>
  static std::unique_ptr<Parser<T>> Create(const std::string& path) {
    gzFile file = gzopen(path.c_str(), "rb");
    if (file == nullptr) {
      throw std::invalid_argument(
          "[bioparser::Parser::Create] error: unable to open file " + path);
    }
    gzbuffer(file, 1 << 20);
    return std::unique_ptr<Parser<T>>(new P<T>(file));
  }

  void Reset() {
    gzseek(file_, 0, SEEK_SET);
    buffer_.clear();
    pos_ = 0;
    eof_ = false;
  }

  // parses records until at least `bytes` bytes of them have been read
  // (all of the file for bytes == -1)
  virtual std::vector<std::unique_ptr<T>> Parse(std::uint64_t bytes,
                                                bool shorten_names = true) = 0;

 protected:
  explicit Parser(gzFile file) : file_(file) {}

  // next line without its terminator; false at the end of the file
  bool ReadLine(std::string* line) {
    line->clear();
    while (true) {
      if (pos_ == buffer_.size()) {
        if (eof_) return !line->empty();
        buffer_.resize(1 << 20);
        const int n = gzread(file_, &buffer_[0], static_cast<unsigned>(buffer_.size()));
        if (n < 0) {
          throw std::invalid_argument("[bioparser::Parser::Parse] error: unable to read file");
        }
        buffer_.resize(static_cast<std::size_t>(n));
        pos_ = 0;
        if (n == 0) {
          eof_ = true;
          return !line->empty();
        }
      }
      const char* begin = buffer_.data() + pos_;
      const char* nl = static_cast<const char*>(std::memchr(begin, '\n', buffer_.size() - pos_));
      if (nl == nullptr) {
        line->append(begin, buffer_.size() - pos_);
        pos_ = buffer_.size();
        continue;
      }
      line->append(begin, static_cast<std::size_t>(nl - begin));
      pos_ += static_cast<std::size_t>(nl - begin) + 1;
      if (!line->empty() && line->back() == '\r') line->pop_back();
      return true;
    }
  }

  static std::uint32_t NameLength(const std::string& header, bool shorten) {
    // header[0] is '>' or '@'
    std::size_t end = header.size();
    if (shorten) {
      for (std::size_t i = 1; i < header.size(); ++i) {
        if (header[i] == ' ' || header[i] == '\t') {
          end = i;
          break;
        }
      }
    }
    return static_cast<std::uint32_t>(end - 1);
  }

  static void StripSpaces(std::string* s) {
    std::size_t k = 0;
    for (char c : *s) {
      if (c != ' ' && c != '\t' && c != '\r') (*s)[k++] = c;
    }
    s->resize(k);
  }

  gzFile file_;
  std::string buffer_;
  std::size_t pos_ = 0;
  bool eof_ = false;
  std::string pending_;  // a header line read ahead of its record
  bool has_pending_ = false;
};

}  // namespace bioparser

#endif  // BIOPARSER_PARSER_HPP_
