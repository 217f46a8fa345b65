module github.com/solarlune/resolv

go 1.20
