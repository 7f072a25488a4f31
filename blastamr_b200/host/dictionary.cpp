#include "dictionary.H"

#include <cctype>
#include <cstdlib>

namespace amr {

class dictParser {
public:
    explicit dictParser(const std::string &t) : s(t) {}
    dictionary parseTop() {
        dictionary d;
        parseBody(d, false);
        return d;
    }

private:
    const std::string &s;
    size_t i = 0;

    void skipWs() {
        while (i < s.size()) {
            if (std::isspace((unsigned char)s[i])) { i++; continue; }
            if (s.compare(i, 2, "//") == 0) { while (i < s.size() && s[i] != '\n') i++; continue; }
            if (s.compare(i, 2, "/*") == 0) { size_t e = s.find("*/", i + 2); i = (e == std::string::npos) ? s.size() : e + 2; continue; }
            break;
        }
    }
    // one token: word/number, quoted string, punctuation ( ) { } ;, or a #{ ... #} block
    bool next(std::string &tok) {
        skipWs();
        if (i >= s.size()) return false;
        char c = s[i];
        if (s.compare(i, 2, "#{") == 0) {
            size_t e = s.find("#}", i + 2);
            if (e == std::string::npos) throw FatalError("dictionary: unterminated #{ block");
            tok = s.substr(i, e + 2 - i); i = e + 2; return true;
        }
        if (c == '(' || c == ')' || c == '{' || c == '}' || c == ';') { tok = std::string(1, c); i++; return true; }
        if (c == '"') {
            size_t e = s.find('"', i + 1);
            if (e == std::string::npos) throw FatalError("dictionary: unterminated string");
            tok = s.substr(i + 1, e - i - 1); i = e + 1; return true;
        }
        size_t b = i;
        while (i < s.size() && !std::isspace((unsigned char)s[i]) && s[i] != ';' && s[i] != '(' && s[i] != ')' && s[i] != '{' && s[i] != '}') i++;
        tok = s.substr(b, i - b);
        return true;
    }
    void parseBody(dictionary &d, bool nested) {
        std::string key;
        while (next(key)) {
            if (key == "}") { if (!nested) throw FatalError("dictionary: unexpected }"); return; }
            if (key == ";") continue;
            std::string t;
            std::vector<std::string> vals;
            bool isSub = false;
            int depth = 0;                       // `key ( a { k v; } b { ... } );` -- PtrList<entry> values nest
            while (next(t)) {
                if (t == ";" && depth == 0) break;
                if (t == "(" || (t == "{" && !vals.empty())) depth++;
                else if ((t == ")" || t == "}") && depth > 0) depth--;
                if (t == "{" && vals.empty()) {
                    auto sub = std::make_shared<dictionary>();
                    sub->name_ = key;
                    parseBody(*sub, true);
                    if (key != "FoamFile") d.subs_[key] = sub;
                    isSub = true;
                    break;
                }
                vals.push_back(t);
            }
            if (!isSub && key != "FoamFile") d.entries_[key] = vals;
        }
        if (nested) throw FatalError("dictionary: missing }");
    }
};

dictionary dictionary::parse(const std::string &text) { return dictParser(text).parseTop(); }

const dictionary &dictionary::subDict(const word &key) const {
    auto it = subs_.find(key);
    if (it == subs_.end()) throw FatalError("keyword " + key + " is undefined in dictionary \"" + name_ + "\" (sub-dictionary expected)");
    return *it->second;
}
const std::vector<std::string> &dictionary::tokens(const word &key) const {
    auto it = entries_.find(key);
    if (it == entries_.end()) throw FatalError("keyword " + key + " is undefined in dictionary \"" + name_ + "\"");
    return it->second;
}
word dictionary::lookupWord(const word &key) const {
    const auto &t = tokens(key);
    if (t.empty()) throw FatalError("keyword " + key + " has no value");
    return t[0];
}
double dictionary::lookupScalar(const word &key) const {
    const word w = lookupWord(key);
    char *e = nullptr;
    double v = std::strtod(w.c_str(), &e);
    if (e == w.c_str() || *e) throw FatalError("keyword " + key + ": expected a scalar, found " + w);
    return v;
}
long dictionary::lookupLabel(const word &key) const {
    const word w = lookupWord(key);
    char *e = nullptr;
    long v = std::strtol(w.c_str(), &e, 10);
    if (e == w.c_str() || *e) throw FatalError("keyword " + key + ": expected a label, found " + w);
    return v;
}
bool dictionary::lookupBool(const word &key) const {
    const word w = lookupWord(key);      // OpenFOAM Switch
    if (w == "yes" || w == "on" || w == "true" || w == "y" || w == "t" || w == "1") return true;
    if (w == "no" || w == "off" || w == "false" || w == "n" || w == "f" || w == "none" || w == "0") return false;
    throw FatalError("keyword " + key + ": expected a Switch, found " + w);
}
wordList dictionary::lookupWordList(const word &key) const {
    const auto &t = tokens(key);
    wordList out;
    size_t k = 0;
    // optional size prefix: `2(a b)` is tokenised as "2" "(" ...
    if (k < t.size() && t[k] != "(" && k + 1 < t.size() && t[k + 1] == "(") k++;
    if (k >= t.size() || t[k] != "(") throw FatalError("keyword " + key + ": expected a list ( ... )");
    for (k++; k < t.size() && t[k] != ")"; k++) out.push_back(t[k]);
    return out;
}
std::vector<std::pair<word, dictionary>> dictionary::lookupEntryList(const word &key) const {
    const auto &t = tokens(key);
    std::vector<std::pair<word, dictionary>> out;
    size_t k = 0;
    if (k < t.size() && t[k] != "(" && k + 1 < t.size() && t[k + 1] == "(") k++;
    if (k >= t.size() || t[k] != "(") throw FatalError("keyword " + key + ": expected a list of entries ( name { ... } ... )");
    k++;
    while (k < t.size() && t[k] != ")") {
        const word name = t[k++];
        if (k >= t.size() || t[k] != "{") throw FatalError("keyword " + key + ": entry " + name + " is not a dictionary");
        int depth = 1;
        std::string body;
        for (k++; k < t.size() && depth > 0; k++) {
            if (t[k] == "{") depth++;
            else if (t[k] == "}") { if (--depth == 0) break; }
            body += t[k];
            body += ' ';
        }
        if (depth != 0) throw FatalError("keyword " + key + ": missing } in entry " + name);
        k++;
        dictionary d = dictionary::parse(body);
        d.name_ = name;
        out.emplace_back(name, d);
    }
    return out;
}
std::vector<word> dictionary::toc() const {
    std::vector<word> out;
    for (auto &e : entries_) out.push_back(e.first);
    for (auto &e : subs_) out.push_back(e.first);
    return out;
}

}  // namespace amr
