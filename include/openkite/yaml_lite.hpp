// yaml_lite.hpp -- ~100-line reader for the flat two-level YAML schema of the kite parameter files
// (top-level scalars and "section:" blocks of "key: value" pairs, '#' comments).  yaml-cpp, which the
// reference uses (kite.cpp:10), is a third-party dependency that is not available here.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>

namespace yaml_lite {

class Document {
public:
    std::map<std::string, std::map<std::string, std::string>> sections;   // "" = top level

    bool has(const std::string& sec, const std::string& key) const {
        auto s = sections.find(sec);
        return s != sections.end() && s->second.count(key) > 0;
    }
    std::string str(const std::string& sec, const std::string& key) const {
        if (!has(sec, key)) throw std::runtime_error("yaml: missing key '" + (sec.empty() ? key : sec + "." + key) + "'");
        return sections.at(sec).at(key);
    }
    double num(const std::string& sec, const std::string& key) const {
        const std::string v = str(sec, key);
        char* end = nullptr;
        double d = std::strtod(v.c_str(), &end);
        if (end == v.c_str() || *end != '\0') throw std::runtime_error("yaml: key '" + sec + "." + key + "' is not a number: '" + v + "'");
        return d;
    }
    double num_or(const std::string& sec, const std::string& key, double dflt) const { return has(sec, key) ? num(sec, key) : dflt; }
};

inline std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

inline Document parse(std::istream& in) {
    Document doc;
    std::string line, section;
    int lineno = 0;
    while (std::getline(in, line)) {
        ++lineno;
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (trim(line).empty()) continue;
        const bool indented = line[0] == ' ' || line[0] == '\t';
        size_t colon = line.find(':');
        if (colon == std::string::npos) throw std::runtime_error("yaml: line " + std::to_string(lineno) + ": expected 'key: value'");
        std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (val.size() >= 2 && ((val.front() == '"' && val.back() == '"') || (val.front() == '\'' && val.back() == '\''))) val = val.substr(1, val.size() - 2);
        if (!indented) {
            if (val.empty()) { section = key; doc.sections[section]; }
            else { section.clear(); doc.sections[""][key] = val; }
        } else {
            if (section.empty()) throw std::runtime_error("yaml: line " + std::to_string(lineno) + ": indented key outside a section");
            doc.sections[section][key] = val;
        }
    }
    return doc;
}

inline Document load_file(const std::string& filename) {
    std::ifstream f(filename);
    if (!f) throw std::runtime_error("yaml: cannot open '" + filename + "'");
    return parse(f);
}

}  // namespace yaml_lite
