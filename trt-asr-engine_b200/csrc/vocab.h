// vocab.h -- the token table of a model directory (vocab.txt, one SentencePiece piece per line, line index == token id) and the two
// text-side operations the runtime needs: ids -> text, and the "punctuation-only piece" predicate of the leading-punctuation
// suppression.  Semantics of /root/reference/cpp/src/tokenizer.cpp:9-84 (Tokenizer::decode / is_punct_only); no GPU involved, so the
// C ABI exposes it on its own (pkb_vocab_*) and the CPU test suite checks it against the reference's golden cases.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace pkb {

class Vocab {
 public:
  Vocab() = default;
  explicit Vocab(const std::string& path);      // throws std::runtime_error (missing / empty file)
  int size() const { return (int)pieces_.size(); }
  const std::string& piece(int id) const;       // "" when out of range
  int find(const std::string& piece) const;     // first id with that text, -1 if none
  bool is_punct_only(int id) const { return id >= 0 && id < size() && (flags_[id] & kPunctOnly); }
  // SentencePiece-style join: specials (<...>) and out-of-range ids are skipped, a piece starting with U+2581 begins a new word,
  // leading spaces are trimmed
  std::string decode(const int* ids, size_t n) const;
  std::string decode(const std::vector<int>& ids) const { return decode(ids.data(), ids.size()); }
  // bit i of word i/32 set <=> piece i is punctuation-only (the device-side table of the decode loop)
  std::vector<uint32_t> punct_bitmap(int n_bits) const;

 private:
  enum : uint8_t { kSpecial = 1, kWordStart = 2, kPunctOnly = 4 };
  std::vector<std::string> pieces_;
  std::vector<uint8_t> flags_;
};

}  // namespace pkb
