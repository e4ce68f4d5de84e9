// Minimal stand-in for the reference's src/models/tokenizer.h:138-347 (CPU-side text <-> ids; NOT on the GPU hot path, provided so
// that examples/cpp/context_decoder_example.cpp compiles unchanged).  Same public calls: Initialize(path), Encode(text),
// Decode(ids).  Vocabulary file layout as the reference reads it (tokenizer.h:138-166): int32 version; if version >= 1 an int32
// count of (string key, string value) pairs, strings = int32 length + bytes; int32 vocab size; per token: int32 length, `length`
// int32 code units (one per byte), int32 id, float32 score.  Encoding: SentencePiece-style greedy merges -- start from single
// bytes of the text with spaces mapped to U+2581, repeatedly join the adjacent pair whose concatenation is a vocabulary entry
// with the highest score.
#pragma once

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

class Tokenizer {
public:
    std::unordered_map<std::string, int> stringToTokenDict;
    std::unordered_map<int, std::string> tokenToStringDict;
    std::unordered_map<std::string, float> scores;

    void Clear() {
        stringToTokenDict.clear();
        tokenToStringDict.clear();
        scores.clear();
    }
    void Insert(const std::string &s, int tokenId, float score = 1.0f) {
        stringToTokenDict[s] = tokenId;
        tokenToStringDict[tokenId] = s;
        scores[s] = score;
    }
    void Initialize(std::string file) {
        std::ifstream in(file, std::ios::binary);
        if (!in.is_open()) {
            std::printf("tokenizer file %s cannot be opened\n", file.c_str());
            return;
        }
        auto read_int = [&]() { int32_t v = 0; in.read(reinterpret_cast<char *>(&v), 4); return (int)v; };
        auto read_float = [&]() { float v = 0; in.read(reinterpret_cast<char *>(&v), 4); return v; };
        auto read_string = [&]() { int n = read_int(); std::string s(n > 0 ? n : 0, '\0'); if (n > 0) in.read(&s[0], n); return s; };
        const int version = read_int();
        if (version >= 1) {
            const int n = read_int();
            for (int i = 0; i < n; ++i) {
                read_string();
                read_string();
            }
        }
        const int vocab = read_int();
        for (int i = 0; i < vocab && in.good(); ++i) {
            const int len = read_int();
            std::string x;
            for (int j = 0; j < len; ++j) x += (char)read_int();
            const int id = read_int();
            const float score = read_float();
            Insert(x, id, score);
        }
    }
    std::vector<int> Encode(const std::string &ori) {
        const std::string blank = "\xE2\x96\x81";
        std::string s = blank;
        for (size_t i = 0; i < ori.size(); ++i) {
            if (ori[i] == ' ') {
                if (i != 0 && ori[i - 1] != ' ') s += blank;
            } else {
                s += ori[i];
            }
        }
        // pieces start as single UTF-8 characters
        std::vector<std::string> pieces;
        for (size_t i = 0; i < s.size();) {
            const unsigned char c = (unsigned char)s[i];
            const size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
            pieces.push_back(s.substr(i, n));
            i += n;
        }
        for (;;) {
            int best = -1;
            float best_score = -1e30f;
            for (size_t i = 0; i + 1 < pieces.size(); ++i) {
                auto it = scores.find(pieces[i] + pieces[i + 1]);
                if (it != scores.end() && it->second > best_score) best_score = it->second, best = (int)i;
            }
            if (best < 0) break;
            pieces[best] += pieces[best + 1];
            pieces.erase(pieces.begin() + best + 1);
        }
        std::vector<int> ids;
        for (const std::string &p : pieces) {
            auto it = stringToTokenDict.find(p);
            if (it != stringToTokenDict.end()) {
                ids.push_back(it->second);
            } else {  // byte fallback: "<0xNN>" entries
                for (unsigned char c : p) {
                    char buf[8];
                    std::snprintf(buf, sizeof(buf), "<0x%02X>", c);
                    auto bt = stringToTokenDict.find(buf);
                    if (bt != stringToTokenDict.end()) ids.push_back(bt->second);
                }
            }
        }
        return ids;
    }
    std::string Decode(const std::vector<int> &ids) {
        std::string out;
        for (int id : ids) {
            auto it = tokenToStringDict.find(id);
            if (it == tokenToStringDict.end()) continue;
            std::string t = it->second;
            for (size_t p; (p = t.find("\xE2\x96\x81")) != std::string::npos;) t.replace(p, 3, " ");
            out += t;
        }
        return out;
    }
    std::string Decode(int id) { return Decode(std::vector<int>{id}); }
};
