// Replays a list of texts / id lists through a Tokenizer and prints the results, one line each.  Built twice from this same source:
// against the reference's header where it lies (oracle/Makefile, target ref_tokenizer -> oracle/_ref/tokenizer_ref: the checker) and
// against the shim's header (__graft_entry__.build() -> shim/_own_programs/tokenizer_shim: the product).  Test infrastructure.
//   usage: tokenizer_driver <vocabulary file> < cases        case = "T <text>"  -> "E id id ..." and "D <hex of Decode(Encode(text))>"
//                                                              "I id id ..."  -> "D <hex of Decode(ids)>"
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include TOKENIZER_HEADER

static void print_hex(const std::string &s) {
    std::printf("D ");
    for (unsigned char c : s) std::printf("%02x", c);
    std::printf("\n");
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    Tokenizer tok;
    tok.Initialize(argv[1]);
    std::string line;
    while (std::getline(std::cin, line)) {
        if (line.size() < 2) continue;
        const std::string body = line.substr(2);
        if (line[0] == 'T') {
            std::vector<int> ids = tok.Encode(body);
            std::printf("E");
            for (int id : ids) std::printf(" %d", id);
            std::printf("\n");
            print_hex(tok.Decode(ids));
        } else if (line[0] == 'I') {
            std::istringstream ss(body);
            std::vector<int> ids;
            for (int v; ss >> v;) ids.push_back(v);
            print_hex(tok.Decode(ids));
        }
    }
    return 0;
}
