module resolvparity

go 1.20

// Point this at the reference checkout being validated:
//   go mod edit -replace github.com/solarlune/resolv=/path/to/resolv && go mod tidy
require github.com/solarlune/resolv v0.8.0
