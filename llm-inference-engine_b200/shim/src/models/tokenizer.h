// Drop-in for the reference's src/models/tokenizer.h:57-347 (CPU-side text <-> ids; not on the GPU hot path -- SURVEY.md 8f rank 3).
// Same public calls and members: Insert(text, id, score), Initialize(path), Encode(text), Decode(ids), DecodeTokens(ids),
// stringToTokenDict, tokenToStringDict.  Results are identical to the reference's, id for id, including its quirks -- checked against
// golden vectors produced by the reference's own header (tests/golden/tokenizer_golden.json, tests/test_tokenizer.py):
//
//   * vocabulary file (tokenizer.h:138-166): int32 version; if version >= 1 an int32 count of (key, value) string pairs (string = int32
//     length + bytes); int32 vocabulary size; per entry int32 length, `length` int32 code units (one per byte), int32 id, float32 score;
//   * Encode (tokenizer.h:184-298): a U+2581 mark in front (not when the text starts with "<FLM_FIX_TOKEN_"), every run of spaces after
//     the first character becomes one mark, leading spaces vanish; "<FLM_FIX_TOKEN_123>" yields the literal id 123; the rest is cut into
//     single BYTES and merged bottom-up: the adjacent pair whose concatenation is a PREFIX of some vocabulary entry with the highest
//     score merges first (ties: leftmost), where a prefix that is not itself an entry scores 0 and carries id 0 -- the reference's trie
//     nodes are value-initialised, so its "-999999 = no token" tests never fire (tokenizer.h:60-64,177-179,229-236).  Bytes that start
//     no entry fall back to the "<0xNN>" entries;
//   * Decode (tokenizer.h:302-347): "<0xNN>" entries become the byte, "<n>" a newline, "<|tab|>" a tab, every mark a space, and a text
//     containing "<|blank_" collapses to atoi(text[8 .. size-2)) spaces.
//
// Implementation: no trie -- one hash map for the entries and one hash set of all their prefixes answer the same two questions
// ("is this string a path of the trie?", "which id / score sits at its end?").
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <queue>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

class Tokenizer {
public:
    std::unordered_map<std::string, int> stringToTokenDict;
    std::unordered_map<int, std::string> tokenToStringDict;

    void Clear() {
        stringToTokenDict.clear();
        tokenToStringDict.clear();
        entry_.clear();
        prefix_.clear();
    }

    void Insert(const std::string &text, int tokenId, float score = 1.0f) {
        for (size_t n = 1; n <= text.size(); ++n) prefix_.insert(text.substr(0, n));
        entry_[text] = Entry{tokenId, score};
        tokenToStringDict[tokenId] = text;
        stringToTokenDict[text] = tokenId;
    }

    void Initialize(std::string file) {
        std::ifstream in(file, std::ios::binary);
        if (!in.is_open()) {
            std::printf("tokenizer file %s cannot be opened\n", file.c_str());
            return;
        }
        auto i32 = [&in]() {
            int32_t v = 0;
            in.read(reinterpret_cast<char *>(&v), 4);
            return (int)v;
        };
        auto f32 = [&in]() {
            float v = 0.0f;
            in.read(reinterpret_cast<char *>(&v), 4);
            return v;
        };
        auto skip_string = [&]() {
            const int n = i32();
            if (n > 0) in.ignore(n);
        };
        if (i32() >= 1) {  // version >= 1: a key-value table comes first
            for (int pairs = i32(); pairs > 0 && in.good(); --pairs) {
                skip_string();
                skip_string();
            }
        }
        for (int left = i32(); left > 0 && in.good(); --left) {
            std::string text;
            for (int units = i32(); units > 0; --units) text += (char)i32();
            const int id = i32();
            const float score = f32();
            Insert(text, id, score);
        }
    }

    std::vector<int> Encode(const std::string &ori) {
        static const std::string kMark = "\xE2\x96\x81", kFix = "<FLM_FIX_TOKEN_";
        // ---- normalise the spaces
        std::string s = (ori.size() > kFix.size() && ori.compare(0, kFix.size(), kFix) == 0) ? std::string() : kMark;
        for (size_t i = 0; i < ori.size(); ++i) {
            if (ori[i] != ' ') s += ori[i];
            else if (i > 0 && ori[i - 1] != ' ') s += kMark;
        }
        // ---- cut into pieces: literal-id markers, single bytes that start some entry, and unknown bytes
        std::vector<Piece> pc;
        for (size_t i = 0; i < s.size(); ++i) {
            if (i + kFix.size() < s.size() && s.compare(i, kFix.size(), kFix) == 0) {
                size_t j = i + kFix.size();
                int id = 0;
                while (j < s.size() && s[j] >= '0' && s[j] <= '9') id = id * 10 + (s[j++] - '0');
                pc.push_back(Piece{std::string(), kLiteral, id, (unsigned char)(j < s.size() ? s[j] : 0)});
                i = j;  // the character after the digits (the '>') is consumed with the marker
                continue;
            }
            const std::string one(1, s[i]);
            if (prefix_.count(one)) pc.push_back(Piece{one, kText, 0, 0});
            else pc.push_back(Piece{std::string(), kUnknown, 0, (unsigned char)s[i]});
        }
        if (pc.empty()) return {};
        const int n = (int)pc.size();
        std::vector<int> prev(n), next(n);
        for (int i = 0; i < n; ++i) prev[i] = i - 1, next[i] = i + 1;
        next[n - 1] = -1;
        // ---- bottom-up merges, best score first, leftmost on ties; an entry of the queue is stale once either side changed length
        struct Cand {
            float score;
            int l, r;
            size_t size;
        };
        auto worse = [](const Cand &a, const Cand &b) { return a.score < b.score || (a.score == b.score && a.l > b.l); };
        std::priority_queue<Cand, std::vector<Cand>, decltype(worse)> queue(worse);
        auto offer = [&](int l, int r) {
            if (l < 0 || r < 0 || pc[l].text.empty() || pc[r].text.empty()) return;
            const std::string joined = pc[l].text + pc[r].text;
            if (!prefix_.count(joined)) return;
            queue.push(Cand{lookup(joined).score, l, r, joined.size()});
        };
        for (int i = 1; i < n; ++i) offer(i - 1, i);
        while (!queue.empty()) {
            const Cand c = queue.top();
            queue.pop();
            if (pc[c.l].text.empty() || pc[c.r].text.empty() || pc[c.l].text.size() + pc[c.r].text.size() != c.size) continue;
            pc[c.l].text += pc[c.r].text;
            pc[c.r].text.clear();
            pc[c.r].kind = kMerged;
            next[c.l] = next[c.r];
            if (next[c.r] >= 0) prev[next[c.r]] = c.l;
            offer(prev[c.l], c.l);
            offer(c.l, next[c.l]);
        }
        // ---- ids
        std::vector<int> ids;
        for (const Piece &p : pc) {
            if (p.kind == kText) {
                ids.push_back(lookup(p.text).id);  // a prefix that is no entry reads id 0, as in the reference
            } else if (p.kind == kLiteral) {
                ids.push_back(p.literal);
            } else if (p.kind == kUnknown) {
                char name[8];
                std::snprintf(name, sizeof(name), "<0x%02X>", p.byte);
                auto it = stringToTokenDict.find(name);
                if (it != stringToTokenDict.end()) ids.push_back(it->second);
            }
        }
        return ids;
    }

    std::string Decode(const std::vector<int> &ids) { return DecodeTokens(ids); }
    std::string Decode(int id) { return DecodeTokens(std::vector<int>{id}); }

    std::string DecodeTokens(const std::vector<int> &ids) {
        static const std::string kMark = "\xE2\x96\x81";
        std::string out;
        for (int id : ids) {
            const std::string &t = tokenToStringDict[id];  // an unknown id adds an empty entry, as the reference's operator[] does
            if (t.size() == 6 && t.compare(0, 3, "<0x") == 0 && t[5] == '>') {
                auto hex = [](char ch) { return ch >= '0' && ch <= '9' ? ch - '0' : ch - 'A' + 10; };
                out += (char)(hex(t[3]) * 16 + hex(t[4]));
            } else if (t == "<n>") {
                out += '\n';
            } else if (t == "<|tab|>") {
                out += '\t';
            } else {
                out += t;
            }
        }
        for (size_t at; (at = out.find(kMark)) != std::string::npos;) out.replace(at, kMark.size(), " ");
        if (out.find("<|blank_") != std::string::npos) return std::string((size_t)std::atoi(out.substr(8, out.size() - 10).c_str()), ' ');
        return out;
    }

private:
    struct Entry {
        int id;
        float score;
    };
    enum Kind { kText, kLiteral, kUnknown, kMerged };
    struct Piece {
        std::string text;  // non-empty only for kText pieces that are still alive
        Kind kind;
        int literal;
        unsigned char byte;
    };
    std::unordered_map<std::string, Entry> entry_;
    std::unordered_set<std::string> prefix_;

    Entry lookup(const std::string &text) const {
        auto it = entry_.find(text);
        return it == entry_.end() ? Entry{0, 0.0f} : it->second;
    }
};
